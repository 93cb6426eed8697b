// ffb_api.cu -- context, stream plumbing and the extern "C" ABI declared in include/ffb.h.
//
// One context = one GPU = one host thread.  Frames flow
//   host (pageable | pinned) --cudaMemcpyAsync on s_copy, double-buffered--> device u8 staging
//   --k_pyramid_pow2 | k_pyramid_level / k_polyexp--> per-frame expansion slots (ring of S = B+2 frames; each frame
//     is expanded once and used as `next` of pair j-1 and `prev` of pair j: streaming mode)
//   --k_flow_iter x3 per level (flow up-sampling fused, k_upsample_flow otherwise)--> final-flow ring (>= B+8 pairs)
//   --k_divmag / k_phase1_finish--> per-pair centre, value, mean magnitude, cut flag (device arrays)
//   --k_smooth_centers / k_radial / k_radial_finish (lagging 6 pairs)--> per-pair scalar
// and only the 1-D per-pair results are copied back at ffb_bracket_finish.
#include "../../include/ffb.h"
#include "ffb_kernels.cuh"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits>
#include <set>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_create_error;

const char* const kKernelNames[FFB_K_COUNT] = {"k_pyramid_level", "k_polyexp",  "k_upsample_flow", "k_flow_iter",
                                               "k_divmag",        "k_radial",   "small",           "k_preprocess"};

// ------------------------------------------------------------------ host-side constant builders
int cv_round(double v) { return (int)nearbyint(v); }   // round-half-even in the default FP mode

struct LevelPlan {
    int n;
    int k[FFB_MAX_LEVELS], w[FFB_MAX_LEVELS], h[FFB_MAX_LEVELS], ksize[FFB_MAX_LEVELS];
    double sigma[FFB_MAX_LEVELS];
};

// calcOpticalFlowFarneback level geometry for pyr_scale 0.5, levels 3.
LevelPlan make_plan(int W, int H) {
    LevelPlan p;
    int k = 0;
    double scale = 1.0;
    while (k < 3) {
        scale *= 0.5;
        if (W * scale < 32 || H * scale < 32) break;
        ++k;
    }
    p.n = k + 1;
    for (int i = 0; i < p.n; ++i) {
        const int kk = k - i;   // coarsest first
        double s = 1.0;
        for (int j = 0; j < kk; ++j) s *= 0.5;
        const double sigma = (1.0 / s - 1.0) * 0.5;
        int ks = cv_round(sigma * 5) | 1;
        if (ks < 3) ks = 3;
        p.k[i] = kk;
        p.w[i] = cv_round(W * s);
        p.h[i] = cv_round(H * s);
        p.ksize[i] = ks;
        p.sigma[i] = sigma;
    }
    return p;
}

// cv::getGaussianKernel(ksize, sigma, CV_32F), half kernel.
FfbTaps make_taps(int ksize, double sigma) {
    FfbTaps t;
    memset(&t, 0, sizeof(t));
    t.r = ksize / 2;
    if (sigma <= 0 && ksize == 3) {
        t.k[0] = 0.5f;
        t.k[1] = 0.25f;
        return t;
    }
    if (sigma <= 0) sigma = ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8;
    std::vector<double> v(ksize);
    double sum = 0;
    for (int i = 0; i < ksize; ++i) {
        const double x = i - (ksize - 1) * 0.5;
        v[i] = exp(-0.5 / (sigma * sigma) * x * x);
        sum += v[i];
    }
    for (int i = 0; i <= t.r; ++i) t.k[i] = (float)(v[t.r + i] / sum);
    return t;
}

// cv::resize INTER_LINEAR per-axis table.
void make_linear_table(int dst_n, int src_n, std::vector<int>& idx, std::vector<float>& alpha) {
    idx.resize(dst_n);
    alpha.resize(dst_n);
    const double scale = (double)src_n / dst_n;
    for (int d = 0; d < dst_n; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int i0 = (int)floorf(f);
        f -= (float)i0;
        if (i0 < 0) { i0 = 0; f = 0.f; }
        if (i0 >= src_n - 1) { i0 = src_n - 1; f = 0.f; }
        idx[d] = i0;
        alpha[d] = f;
    }
}

// FarnebackPrepareGaussian(n = 5, sigma = 1.2): taps in float32, inverse moments in double.
void invert6(double G[6][6], double inv[6][6]) {
    double a[6][12];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) { a[i][j] = G[i][j]; a[i][j + 6] = i == j ? 1.0 : 0.0; }
    for (int c = 0; c < 6; ++c) {
        int piv = c;
        for (int r = c + 1; r < 6; ++r) if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
        for (int j = 0; j < 12; ++j) { double t = a[c][j]; a[c][j] = a[piv][j]; a[piv][j] = t; }
        const double d = a[c][c];
        for (int j = 0; j < 12; ++j) a[c][j] /= d;
        for (int r = 0; r < 6; ++r) {
            if (r == c) continue;
            const double f = a[r][c];
            for (int j = 0; j < 12; ++j) a[r][j] -= f * a[c][j];
        }
    }
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) inv[i][j] = a[i][j + 6];
}

FfbPolyConsts make_poly_consts() {
    const int n = FFB_POLY_N;
    const double sigma = 1.2;
    float g[2 * FFB_POLY_N + 1], xg[2 * FFB_POLY_N + 1], xxg[2 * FFB_POLY_N + 1];
    double s = 0;
    for (int x = -n; x <= n; ++x) { g[x + n] = (float)exp(-x * x / (2 * sigma * sigma)); s += g[x + n]; }
    s = 1.0 / s;
    for (int x = -n; x <= n; ++x) {
        g[x + n] = (float)(g[x + n] * s);
        xg[x + n] = (float)(x * g[x + n]);
        xxg[x + n] = (float)(x * x * g[x + n]);
    }
    double G[6][6];
    memset(G, 0, sizeof(G));
    for (int y = -n; y <= n; ++y)
        for (int x = -n; x <= n; ++x) {
            const float gg = g[y + n] * g[x + n];
            G[0][0] += gg;
            G[1][1] += gg * (float)(x * x);
            G[3][3] += gg * (float)(x * x * x * x);
            G[5][5] += gg * (float)(x * x * y * y);
        }
    G[2][2] = G[0][3] = G[0][4] = G[3][0] = G[4][0] = G[1][1];
    G[4][4] = G[3][3];
    G[3][4] = G[4][3] = G[5][5];
    double inv[6][6];
    invert6(G, inv);
    FfbPolyConsts c;
    for (int i = 0; i <= n; ++i) { c.g[i] = g[n + i]; c.xg[i] = xg[n + i]; c.xxg[i] = xxg[n + i]; }
    c.ig11 = (float)inv[1][1];
    c.ig03 = (float)inv[0][3];
    c.ig33 = (float)inv[3][3];
    c.ig55 = (float)inv[5][5];
    return c;
}

// ------------------------------------------------------------------ context
struct Level {
    int k = 0, w = 0, h = 0, ksize = 3;
    double sigma = 0;
    int rp = 0;            // R / I row pitch (floats)
    size_t plane = 0;      // rp * h
    int fp = 0;            // flow row pitch (float2)
    size_t r_off = 0;      // float offset of this level inside a frame's expansion slot
    FfbTaps taps;
    int RW = 0, RH = 0;    // source window bound of one 32x8 output tile
    int *xi = nullptr, *yi = nullptr, *uxi = nullptr, *uyi = nullptr;
    float *xa = nullptr, *ya = nullptr, *uxa = nullptr, *uya = nullptr;
    float* I = nullptr;    // [B][h][rp]
    float2 *fA = nullptr, *fB = nullptr;   // [B][h][fp]
};

struct ProfRec { int kid; int level; double bytes; cudaEvent_t e0, e1; };
typedef int PtrKindTag;

}  // namespace

struct ffb_ctx {
    int device = 0;
    std::string err;
    cudaStream_t s_comp = nullptr, s_copy = nullptr;
    cudaStream_t s_aux[3] = {nullptr, nullptr, nullptr};   // extra streams for the sliced flow phase
    cudaStream_t launch_stream = nullptr;              // stream of the flow-iteration launch in flight
    cudaStream_t aux_stream = nullptr;                 // stream of the divergence / magnitude launch in flight
    cudaStream_t s_time = nullptr;                     // only carries the end-of-flow-phase timing event
    cudaEvent_t ev_flow_end[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_chain = nullptr;                    // "everything queued on s_comp so far": ffb_chain_after
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_expand[2] = {nullptr, nullptr};
    FfbPolyConsts poly;
    // geometry
    int W = 0, H = 0, B = 0, maxPairs = 0, nlev = 0, S = 0, ring_n = 0, fp0 = 0;
    int Bcap = 0, pairsCap = 0;                 // allocated capacities (B / maxPairs are the limits of the last ffb_configure)
    int64_t n_dev_alloc = 0, n_host_alloc = 0;  // cudaMalloc / cudaHostAlloc calls made for this context so far
    bool pyr_fast = false;
    Level lev[FFB_MAX_LEVELS];
    float* R = nullptr;
    size_t r_slot_floats = 0;
    long long seg_frame_px = 0;      // pixels of the frame whose levels are being iterated (segmentation rule)
    float2* ring = nullptr;
    size_t ring_stride = 0;   // float2 per ring element
    uint8_t* d_u8[2] = {nullptr, nullptr};
    uint8_t* h_pin[2] = {nullptr, nullptr};
    // per-bracket results (device) + partials
    int *d_cx = nullptr, *d_cy = nullptr;
    int *d_cx_ext = nullptr, *d_cy_ext = nullptr;   // raw centres handed in by ffb_bracket_radial (shard + 6 each side)
    float *d_val = nullptr, *d_mm = nullptr;
    unsigned char* d_cut = nullptr;
    double *d_centers = nullptr, *d_scalar = nullptr;
    unsigned long long* d_pkey = nullptr;
    double *d_psum = nullptr, *d_rpart = nullptr;
    int div_nblk = 0, div_rpb = 0, div_gx = 0, div_gy = 0, rad_gx = 0, rad_gy = 0, rad_rpb = 0;
    char* h_res = nullptr;    // pinned result staging
    // frame pre-processing (row N2): source geometry, tables, chunked colour staging
    // source size -> resize target (TW x TH) -> kept window [y0, y0+OH) x [x0, x0+OW) of the target
    struct PrePlan { int W = 0, H = 0, TW = 0, TH = 0, x0 = 0, y0 = 0, OW = 0, OH = 0;
                     bool operator==(const PrePlan& o) const {
                         return W == o.W && H == o.H && TW == o.TW && TH == o.TH && x0 == o.x0 && y0 == o.y0 && OW == o.OW && OH == o.OH; } } pre;
    int *pre_xt = nullptr, *pre_yt = nullptr;
    uint8_t* d_color[2] = {nullptr, nullptr};
    uint8_t* h_color[2] = {nullptr, nullptr};
    cudaEvent_t ev_ch2d[2] = {nullptr, nullptr}, ev_pre[2] = {nullptr, nullptr};
    int color_no = 0;
    // bracket state
    bool in_bracket = false;
    int pov = 0;
    float thr = 7.f;
    int frames_seen = 0, pairs_done = 0, radial_done = 0, batch_no = 0;
    bool deferred = false;      // shard mode: the radial pass waits for ffb_bracket_radial (external raw centres)
    bool phase1_read = false;   // ffb_bracket_phase1_finish has been called for the open bracket
    int pend = 0;               // pre-processed gray frames waiting in d_u8[batch_no & 1] for a full batch
    PtrKindTag pend_kind = 0;   // source kind of the colour frames they came from (first-batch policy)
    // instrumentation
    bool prof = false;
    std::vector<ProfRec> recs;
    std::vector<cudaEvent_t> ev_pool;
    int64_t k_launches[FFB_K_COUNT] = {0};
    double k_ms[FFB_K_COUNT] = {0}, k_bytes[FFB_K_COUNT] = {0};
    int64_t it_launches[FFB_MAX_LEVELS] = {0};          // k_flow_iter split by level k
    double it_ms[FFB_MAX_LEVELS] = {0}, it_bytes[FFB_MAX_LEVELS] = {0};
    // cudaFuncSetAttribute is per device: remember per context which kernels already had their
    // dynamic shared-memory limit raised (a process may own contexts on several GPUs)
    size_t pyr_smem_max = 48 * 1024;
    bool attr_poly = false, attr_pyr2 = false;
    std::set<const void*> attr_iter;                     // k_flow_iter instantiations already configured
    int cur_level = -1;                                  // level k of the flow iteration being launched
    bool phase_timing = false;                           // sliced flow phase: time the phase, not the launches
    bool sliced_divmag = false;                          // flow_pairs already launched k_divmag per slice
    cudaEvent_t phase_rec_e1 = nullptr;
    bool prof_open = false;                              // a per-launch record is waiting for its end event
    int flow_streams = 2;
    int copy_threads = 4;                                // host threads that stage pageable frames into pinned memory
    int64_t launches = 0;
    cudaEvent_t timers[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

namespace {

int fail(ffb_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CK(c, call)                                                                                   \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail((c), FFB_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define CKL(c)                                                                                        \
    do {                                                                                              \
        cudaError_t e_ = cudaGetLastError();                                                          \
        if (e_ != cudaSuccess)                                                                        \
            return fail((c), FFB_E_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define TRY(expr)                   \
    do {                            \
        int rc_ = (expr);           \
        if (rc_ != FFB_OK) return rc_; \
    } while (0)

template <class T>
int dev_alloc(ffb_ctx* c, T** p, size_t count) {
    void* v = nullptr;
    cudaError_t e = cudaMalloc(&v, count * sizeof(T) + 256);   // +256: vector loads may touch a row's tail
    if (e != cudaSuccess) return fail(c, FFB_E_NOMEM, "cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
    if (c) c->n_dev_alloc++;
    *p = (T*)v;
    return FFB_OK;
}
template <class T>
void dev_free(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}
template <class T>
int upload_vec(ffb_ctx* c, T** dptr, const std::vector<T>& v) {
    TRY(dev_alloc(c, dptr, v.size()));
    CK(c, cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return FFB_OK;
}

// ---- profiling helpers
void prof_begin(ffb_ctx* c, int kid, double bytes) {
    c->launches++;
    c->k_launches[kid]++;
    c->k_bytes[kid] += bytes;
    const int lvl = (kid == FFB_K_FLOW_ITER && c->cur_level >= 0 && c->cur_level < FFB_MAX_LEVELS) ? c->cur_level : -1;
    if (lvl >= 0) { c->it_launches[lvl]++; c->it_bytes[lvl] += bytes; }
    if (!c->prof || (kid == FFB_K_FLOW_ITER && c->phase_timing)) return;
    ProfRec r;
    r.kid = kid;
    r.level = lvl;
    r.bytes = bytes;
    for (cudaEvent_t* e : {&r.e0, &r.e1}) {
        if (!c->ev_pool.empty()) { *e = c->ev_pool.back(); c->ev_pool.pop_back(); }
        else cudaEventCreate(e);
    }
    cudaEventRecord(r.e0, kid == FFB_K_FLOW_ITER ? c->launch_stream : (kid == FFB_K_DIVMAG ? c->aux_stream : c->s_comp));
    c->recs.push_back(r);
    c->prof_open = true;
}
void prof_end(ffb_ctx* c) {
    if (!c->prof_open) return;          // prof_begin did not open a per-launch record
    c->prof_open = false;
    const int kid_ = c->recs.back().kid;
    cudaEventRecord(c->recs.back().e1, kid_ == FFB_K_FLOW_ITER ? c->launch_stream : (kid_ == FFB_K_DIVMAG ? c->aux_stream : c->s_comp));
}
// Sliced flow phase: the launch chains of the slices overlap, so per-launch event times would count
// the same device time several times.  The phase is timed as a whole on s_comp instead
// (fork ... join) and attributed to k_flow_iter.
void prof_phase_begin(ffb_ctx* c) {
    if (!c->prof) return;
    ProfRec r;
    r.kid = FFB_K_FLOW_ITER;
    r.level = -1;
    r.bytes = 0;
    for (cudaEvent_t* e : {&r.e0, &r.e1}) {
        if (!c->ev_pool.empty()) { *e = c->ev_pool.back(); c->ev_pool.pop_back(); }
        else cudaEventCreate(e);
    }
    cudaEventRecord(r.e0, c->s_comp);
    c->phase_rec_e1 = r.e1;
    c->recs.push_back(r);
}
// the phase ends when the last slice's last flow iteration does: a side stream waits for every slice's
// end-of-flow event and records the timing event, so that the per-slice reductions that follow on the
// slice streams are not counted
void prof_phase_end(ffb_ctx* c, int nslice) {
    if (!c->prof) return;
    for (int i = 0; i < nslice; ++i) cudaStreamWaitEvent(c->s_time, c->ev_flow_end[i], 0);
    cudaEventRecord(c->phase_rec_e1, c->s_time);
}

void prof_collect(ffb_ctx* c) {
    for (ProfRec& r : c->recs) {
        float ms = 0.f;
        cudaEventSynchronize(r.e1);
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        c->k_ms[r.kid] += ms;
        if (r.level >= 0) c->it_ms[r.level] += ms;
        c->ev_pool.push_back(r.e0);
        c->ev_pool.push_back(r.e1);
    }
    c->recs.clear();
}

// ------------------------------------------------------------------ kernel launchers
constexpr int PYR_TW = 32, PYR_TH = 8;

// Upper bound of the source window (in source pixels) any 32x8 output tile of a level needs.
void pyramid_window(const std::vector<int>& idx, int dst_n, int src_n, int tile, int r, int* bound) {
    int best = 0;
    for (int t0 = 0; t0 < dst_n; t0 += tile) {
        const int t1 = (t0 + tile < dst_n ? t0 + tile : dst_n) - 1;
        int hi = idx[t1] + 1;
        if (hi > src_n - 1) hi = src_n - 1;
        const int ext = (hi + r) - (idx[t0] - r) + 1;
        if (ext > best) best = ext;
    }
    *bound = best;
}

int launch_pyramid(ffb_ctx* c, const uint8_t* src, size_t src_stride, int src_pitch, int W, int H, float* dst,
                   size_t dst_stride, int dp, int w, int h, const int* xi, const float* xa, const int* yi,
                   const float* ya, const FfbTaps& taps, int RW, int RH, int nframes) {
    FfbPyrArgs a;
    a.src = src; a.src_frame_stride = src_stride; a.src_pitch = src_pitch; a.W = W; a.H = H;
    a.dst = dst; a.dst_frame_stride = dst_stride; a.dp = dp; a.w = w; a.h = h;
    a.xi = xi; a.xa = xa; a.yi = yi; a.ya = ya; a.taps = taps;
    a.RWp = RW | 1;                       // odd pitch: column walks of pass 1 stay conflict-free
    a.p_off = ffb_round_up(a.RWp * RH, 4);
    const size_t smem = ((size_t)a.p_off + (size_t)RH * PYR_TW) * sizeof(float);
    if (taps.r >= W || taps.r >= H) return fail(c, FFB_E_INVALID, "frame smaller than the blur radius");
    auto kfn = k_pyramid_level<PYR_TW, PYR_TH>;
    if (smem > c->pyr_smem_max) {
        CK(c, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        c->pyr_smem_max = smem;
    }
    dim3 grid((w + PYR_TW - 1) / PYR_TW, (h + PYR_TH - 1) / PYR_TH, nframes);
    prof_begin(c, FFB_K_PYRAMID, (double)nframes * ((double)W * H + 4.0 * w * h));
    FFB_LAUNCH(kfn, grid, dim3(PYR_TW * PYR_TH), smem, c->s_comp, a);
    prof_end(c);
    CKL(c);
    return FFB_OK;
}

int launch_polyexp(ffb_ctx* c, const float* src, size_t src_stride, int sp, int w, int h, FfbRing dst, size_t plane,
                   int rp, int nframes) {
    FfbPolyArgs a;
    a.src = src; a.src_frame_stride = src_stride; a.sp = sp; a.w = w; a.h = h;
    a.dst = dst; a.plane = plane; a.rp = rp; a.c = c->poly;
    // the kernel stores two pixels (32 bytes) of the float4 image per instruction
    if (rp % 4 != 0 || ((uintptr_t)dst.base | (uintptr_t)dst.stride) % 32 != 0)
        return fail(c, FFB_E_INVALID, "expansion output is not 32-byte aligned (pitch %d, stride %zu)", rp, dst.stride);
    if (!c->attr_poly) {
        CK(c, cudaFuncSetAttribute(k_polyexp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POLY_SMEM));
        c->attr_poly = true;
    }
    // (a variant with packed fp32 arithmetic -- FFMA2 / FADD2 / FMUL2 on row pairs -- was measured in round 2: 4.60 ms
    // against 3.92 ms per 257 1080p frames for this scalar kernel; the packed instructions issue at half rate and the
    // 114 registers they need halve the occupancy.  profiles/r2_other_kernels.txt)
    dim3 grid((w + POLY_OW - 1) / POLY_OW, (h + POLY_ROWS - 1) / POLY_ROWS, nframes);
    prof_begin(c, FFB_K_POLYEXP, (double)nframes * 24.0 * w * h);
    FFB_LAUNCH(k_polyexp, grid, dim3(256), POLY_SMEM, c->s_comp, a);
    prof_end(c);
    CKL(c);
    return FFB_OK;
}

int launch_upsample(ffb_ctx* c, const float2* src, size_t src_stride, int sp, int wc, int hc, float2* dst,
                    size_t dst_stride, int dp, int w, int h, const int* xi, const float* xa, const int* yi,
                    const float* ya, int npairs) {
    FfbUpArgs a;
    a.src = src; a.src_stride = src_stride; a.sp = sp; a.wc = wc; a.hc = hc;
    a.dst = dst; a.dst_stride = dst_stride; a.dp = dp; a.w = w; a.h = h;
    a.xi = xi; a.xa = xa; a.yi = yi; a.ya = ya;
    dim3 grid((w + 31) / 32, (h + 7) / 8, npairs);
    prof_begin(c, FFB_K_UPSAMPLE, (double)npairs * (8.0 * wc * hc + 8.0 * w * h));
    FFB_LAUNCH(k_upsample_flow, grid, dim3(256), 0, c->s_comp, a);
    prof_end(c);
    CKL(c);
    return FFB_OK;
}

// Tunables of the fused iteration kernel.  The defaults are compiled in; the environment overrides exist for
// tuning runs and for the tests that force a variant onto small frames (parsed per launch):
//   FFB_ITER_CFG=NTxUxHO   threads per CTA (= matrix columns per strip) x rows per step x outputs per horizontal task;
//                          compiled in: 256x4x8, 128x2x4, 160x2x4, 96x2x4, 64x2x4 (the round-2 sweeps measured and dropped 192 /
//                          512-thread strips, 8-output tasks on 2-row steps, 4-output tasks on 4-row steps, non-allocating
//                          loads, bulk L2 prefetch, early issue of the second row pair, and a barrier-free variant in which
//                          neighbouring warps synchronise through shared-memory flags: profiles/r2_sweep_flow_iter.txt)
//   FFB_ITER_CFG_K<k>      the same for pyramid level k only
//   FFB_ITER_SH / FFB_ITER_MINSEG   rows per march segment (upper bound) / minimum segments per level
//   FFB_ITER_SWMAX=0       equal-width strips (round 1) instead of full-width strips plus a narrow last one
struct IterCfg { int nt, u, ho, sh, swmax; };
IterCfg iter_cfg() {
    IterCfg c{0, 2, 4, 270, 1};
    if (const char* e = getenv("FFB_ITER_CFG")) {
        int nt = 0, u = 0, ho = 0;
        if (sscanf(e, "%dx%dx%d", &nt, &u, &ho) == 3) { c.nt = nt; c.u = u; c.ho = ho; }
    }
    if (const char* e = getenv("FFB_ITER_SH")) { const int v = atoi(e); if (v >= 16) c.sh = v; }
    if (const char* e = getenv("FFB_ITER_SWMAX")) c.swmax = atoi(e) != 0;
    return c;
}

template <int NT, int U, int MINB, int HO>
int launch_flow_iter_t(ffb_ctx* c, FfbIterArgs a, int npairs, int sh_target, double bytes) {
    const int w = a.w, h = a.h;
    const int sw_max = (NT - 2 * FFB_WIN_R) / HO * HO;
    if (iter_cfg().swmax) {
        // full-width strips; the last one is narrower and its idle warps skip the gather (k_flow_iter `live`)
        a.SW = sw_max;
    } else {
        const int nstrips = (w + sw_max - 1) / sw_max;
        a.SW = ffb_round_up((w + nstrips - 1) / nstrips, HO);
        if (a.SW > sw_max) a.SW = sw_max;
    }
    const int gx = (w + a.SW - 1) / a.SW;
    // Row segments depend on the level geometry only (never on the batch composition), so a pair's
    // result is bit-identical however frames are batched or sharded.
    // at most sh_target rows per segment, at least min_seg segments per level (coarse levels would
    // otherwise be a handful of long, latency-bound marches), never under 32 rows
    const int min_seg_env = getenv("FFB_ITER_MINSEG") ? atoi(getenv("FFB_ITER_MINSEG")) : 0;
    // frames of 1280x720 and more: at least 3 segments on every level (parallelism for their 64-pair batches; 2 .. 4
    // measure the same at 1080p, 5 and more lose); smaller frames come in batches of hundreds, and every segment pays
    // 14 halo rows: 2 (+7 % at 256x256, +3.5 % at 640x360) -- profiles/r1_sweep_segments.txt, r2_sweep_flow_iter.txt.
    // The rule looks at the frame the context is configured for, never at the batch.
    const long long frame_px = c->seg_frame_px > 0 ? c->seg_frame_px : (long long)w * h;      // stage hooks: the level itself
    const int frame_w0 = c->seg_frame_px > 0 ? c->W : w;
    const int min_seg = min_seg_env > 0 ? min_seg_env : (frame_px >= 1280LL * 720 ? 3 : (frame_w0 <= 320 ? 1 : 2));
    int nseg = (h + sh_target - 1) / sh_target;
    if (nseg < min_seg) nseg = min_seg;
    if (nseg > (h + 31) / 32) nseg = (h + 31) / 32;
    if (nseg < 1) nseg = 1;
    a.SH = ffb_round_up((h + nseg - 1) / nseg, U);     // even: see the fused up-sampling in k_flow_iter
    const int gy = (h + a.SH - 1) / a.SH;
    auto kfn = a.up_src ? k_flow_iter<NT, U, MINB, true, HO> : k_flow_iter<NT, U, MINB, false, HO>;
    const size_t smem = ffb_flow_iter_smem<NT, U, HO>();
    if (!c->attr_iter.count((const void*)kfn)) {
        CK(c, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        c->attr_iter.insert((const void*)kfn);
    }
    prof_begin(c, FFB_K_FLOW_ITER, bytes);
    // pair index varies fastest in launch order: the CTAs working on one spatial tile of consecutive
    // pairs are co-resident, so frame j+1's expansion (R1 of pair j, R0 of pair j+1) is read from
    // HBM once and hit in L2 the second time.
    FFB_LAUNCH(kfn, dim3(npairs, gx, gy), dim3(NT), smem, c->launch_stream, a);
    prof_end(c);
    CKL(c);
    return FFB_OK;
}

struct UpSrc {   // coarser-level flow to be up-sampled inside the iteration kernel (A1e fused)
    const float2* src = nullptr; size_t stride = 0; int sp = 0, wc = 0, hc = 0;
};

int launch_flow_iter(ffb_ctx* c, FfbRing R, size_t plane, int rp, int w, int h, const float2* fin, size_t fin_stride,
                     int fip, FfbRing fout, int fop, int npairs, const UpSrc* up = nullptr) {
    FfbIterArgs a;
    a.R = R; a.plane = (int)plane; a.rp = rp; a.w = w; a.h = h;
    a.fin = fin; a.fin_stride = fin_stride; a.fip = fip; a.fout = fout; a.fop = fop;
    a.SW = a.SH = 0;
    a.up_src = nullptr; a.up_stride = 0; a.usp = a.wc = a.hc = 0;
    double bpp = fin ? 56.0 : 48.0;
    if (up && up->src) {
        a.up_src = up->src; a.up_stride = up->stride; a.usp = up->sp; a.wc = up->wc; a.hc = up->hc;
        a.fin = nullptr;
        bpp = 48.0 + 8.0 * ((double)up->wc * up->hc) / ((double)w * h);
    }
    const double bytes = (double)npairs * bpp * w * h;
    // the kernel stores four flow vectors with one 32-byte instruction
    if (fop % 4 != 0 || ((uintptr_t)fout.base | (uintptr_t)fout.stride) % 32 != 0)
        return fail(c, FFB_E_INVALID, "flow output is not 32-byte aligned (pitch %d, stride %zu)", fop, fout.stride);
    const IterCfg k = iter_cfg();
    int nt = k.nt, u = k.u, ho = k.ho;
    if (c->cur_level >= 0 && c->cur_level < FFB_MAX_LEVELS) {      // FFB_ITER_CFG_K<k>: one level only
        char name[24];
        snprintf(name, sizeof(name), "FFB_ITER_CFG_K%d", c->cur_level);
        if (const char* e = getenv(name)) {
            int a1 = 0, a2 = 0, a3 = 0;
            if (sscanf(e, "%dx%dx%d", &a1, &a2, &a3) == 3) { nt = a1; u = a2; ho = a3; }
        }
    }
    if (nt == 0) {
        // Defaults from the round-2 sweeps on a B200 (profiles/r2_sweep_flow_iter.txt), keyed on the frame the context
        // is configured for (never on the batch):
        //   1280x720 and larger   256 threads, 4 rows per step, 8 outputs per horizontal task: 242-column strips carry
        //                         5.5 % halo columns instead of 10.9 %, and the 120 horizontal tasks of a step fill 4 of
        //                         the 8 warps (1080p: +8.4 % over 128x2x4, 4K: +7 %, 720p: +1 %)
        //   up to 320 columns     160 threads x 2 rows x 4 outputs (the reference's 256x256 product mode: 2 strips); its
        //                         64- and 32-column levels take 96- and 64-thread strips (5 / 8 CTAs per SM instead of
        //                         mostly idle 160-thread ones) and a single row segment per level: +3 % together
        //   in between            128 threads x 2 rows x 4 outputs (640x360: +8 % over 160x2x4)
        const long long frame_px = c->seg_frame_px > 0 ? c->seg_frame_px : (long long)w * h;
        const int frame_w = c->seg_frame_px > 0 ? c->W : w;
        if (frame_px >= 1280LL * 720) { nt = 256; u = 4; ho = 8; }
        else if (frame_w <= 320) { nt = w <= 48 ? 64 : (w <= 80 ? 96 : 160); u = 2; ho = 4; }     // by LEVEL width
        else { nt = 128; u = 2; ho = 4; }
    }
    switch (nt * 100 + u * 10 + ho) {
        case 25648: return launch_flow_iter_t<256, 4, 2, 8>(c, a, npairs, k.sh, bytes);
        case 12824: return launch_flow_iter_t<128, 2, 4, 4>(c, a, npairs, k.sh, bytes);
        case 16024: return launch_flow_iter_t<160, 2, 3, 4>(c, a, npairs, k.sh, bytes);
        case 9624:  return launch_flow_iter_t<96, 2, 5, 4>(c, a, npairs, k.sh, bytes);
        case 6424:  return launch_flow_iter_t<64, 2, 8, 4>(c, a, npairs, k.sh, bytes);
        default: break;
    }
    return fail(c, FFB_E_INVALID, "FFB_ITER_CFG: no k_flow_iter variant %dx%dx%d (compiled in: 256x4x8, 128x2x4, 160x2x4, 96x2x4, 64x2x4)", nt, u, ho);
}

// ------------------------------------------------------------------ geometry
// Device / pinned memory comes in four groups with their own lifetimes, so that ffb_configure only
// re-allocates what a request outgrows:
//   frame size   level tables (freed when width or height change, together with everything below)
//   batch        level images, flow ping-pong, expansion ring, frame staging     capacity Bcap frames
//   flow ring    final-flow ring and the reduction partials indexed by ring slot  capacity ring_n pairs
//   pairs        per-pair results and their pinned staging                        capacity pairsCap pairs
void free_batch_group(ffb_ctx* c) {
    for (int l = 0; l < FFB_MAX_LEVELS; ++l) {
        Level& L = c->lev[l];
        dev_free(L.I); dev_free(L.fA); dev_free(L.fB);
    }
    dev_free(c->R);
    for (int b = 0; b < 2; ++b) {
        dev_free(c->d_u8[b]);
        if (c->h_pin[b]) cudaFreeHost(c->h_pin[b]);
        c->h_pin[b] = nullptr;
    }
    c->Bcap = 0;
    c->S = 0;
}
void free_ring_group(ffb_ctx* c) {
    dev_free(c->ring); dev_free(c->d_pkey); dev_free(c->d_psum); dev_free(c->d_rpart);
    c->ring_n = 0;
}
void free_pairs_group(ffb_ctx* c) {
    dev_free(c->d_cx); dev_free(c->d_cy); dev_free(c->d_val); dev_free(c->d_mm); dev_free(c->d_cut);
    dev_free(c->d_centers); dev_free(c->d_scalar); dev_free(c->d_cx_ext); dev_free(c->d_cy_ext);
    if (c->h_res) cudaFreeHost(c->h_res);
    c->h_res = nullptr;
    c->pairsCap = 0;
}
void free_geometry(ffb_ctx* c) {
    free_batch_group(c);
    free_ring_group(c);
    free_pairs_group(c);
    for (int l = 0; l < FFB_MAX_LEVELS; ++l) {
        Level& L = c->lev[l];
        dev_free(L.xi); dev_free(L.yi); dev_free(L.uxi); dev_free(L.uyi);
        dev_free(L.xa); dev_free(L.ya); dev_free(L.uxa); dev_free(L.uya);
        L = Level();
    }
    c->W = c->H = c->B = c->maxPairs = c->nlev = 0;
}

int host_alloc(ffb_ctx* c, void** p, size_t bytes) {
    if (cudaHostAlloc(p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return fail(c, FFB_E_NOMEM, "cudaHostAlloc(%zu) failed", bytes);
    }
    c->n_host_alloc++;
    return FFB_OK;
}

int build_level_tables(ffb_ctx* c, Level& L, int W, int H, int wc, int hc) {
    std::vector<int> xi, yi;
    std::vector<float> xa, ya;
    make_linear_table(L.w, W, xi, xa);
    make_linear_table(L.h, H, yi, ya);
    pyramid_window(xi, L.w, W, PYR_TW, L.taps.r, &L.RW);
    pyramid_window(yi, L.h, H, PYR_TH, L.taps.r, &L.RH);
    TRY(upload_vec(c, &L.xi, xi)); TRY(upload_vec(c, &L.xa, xa));
    TRY(upload_vec(c, &L.yi, yi)); TRY(upload_vec(c, &L.ya, ya));
    if (wc > 0) {
        make_linear_table(L.w, wc, xi, xa);
        make_linear_table(L.h, hc, yi, ya);
        TRY(upload_vec(c, &L.uxi, xi)); TRY(upload_vec(c, &L.uxa, xa));
        TRY(upload_vec(c, &L.uyi, yi)); TRY(upload_vec(c, &L.uya, ya));
    }
    return FFB_OK;
}

bool pyramid_fast_ok(int W, int H, const LevelPlan& p);

int quiesce(ffb_ctx* c) {
    CK(c, cudaStreamSynchronize(c->s_comp));
    CK(c, cudaStreamSynchronize(c->s_copy));
    for (int i = 0; i < 3; ++i) CK(c, cudaStreamSynchronize(c->s_aux[i]));
    return FFB_OK;
}

// ring_pairs: final flows that must stay resident at once (0 = the streaming minimum, batch + 8)
int configure(ffb_ctx* c, int W, int H, int B, int maxPairs, int ring_pairs = 0) {
    // arguments are checked before anything is touched: a refused call leaves the configuration as it was
    if (W < 16 || H < 16 || B < 1 || maxPairs < 1 || ring_pairs < 0 || (double)W * H >= 4294967295.0)
        return fail(c, FFB_E_INVALID, "ffb_configure: bad geometry %dx%d batch %d pairs %d", W, H, B, maxPairs);
    if (c->in_bracket) return fail(c, FFB_E_INVALID, "ffb_configure inside a bracket");
    const int need_ring = ring_pairs > B + 8 ? ring_pairs : B + 8;
    const bool new_frame = W != c->W || H != c->H;
    if (!new_frame && B <= c->Bcap && need_ring <= c->ring_n && maxPairs <= c->pairsCap) {
        c->B = B;                 // logical limits; the capacities stay
        c->maxPairs = maxPairs;
        return FFB_OK;
    }
    TRY(quiesce(c));
    if (new_frame) {
        free_geometry(c);
        const LevelPlan p = make_plan(W, H);
        c->W = W; c->H = H; c->nlev = p.n;
        c->pyr_fast = pyramid_fast_ok(W, H, p);
        size_t off = 0;
        for (int l = 0; l < p.n; ++l) {
            Level& L = c->lev[l];
            L.k = p.k[l]; L.w = p.w[l]; L.h = p.h[l]; L.ksize = p.ksize[l]; L.sigma = p.sigma[l];
            L.rp = ffb_round_up(L.w, 4);
            L.plane = (size_t)L.rp * L.h;
            L.fp = ffb_round_up(L.w, 4);
            L.r_off = off;
            off += 5 * L.plane;
            off = (off + 63) / 64 * 64;
            L.taps = make_taps(L.ksize, L.sigma);
            TRY(build_level_tables(c, L, W, H, l > 0 ? c->lev[l - 1].w : 0, l > 0 ? c->lev[l - 1].h : 0));
        }
        c->r_slot_floats = (off + 127) / 128 * 128;      // 512-byte multiples
        c->fp0 = c->lev[p.n - 1].fp;
        c->ring_stride = (size_t)c->fp0 * H;
        // reduction geometry: fixed per frame size so results do not depend on batch composition
        c->div_rpb = 8;       // whole rows per block: long contiguous reads (a column-marching variant measured slower)
        c->div_gx = (H + c->div_rpb - 1) / c->div_rpb;
        c->div_gy = 1;
        c->div_nblk = c->div_gx;
        c->rad_gx = (W + 255) / 256;
        c->rad_rpb = 32;
        c->rad_gy = (H + c->rad_rpb - 1) / c->rad_rpb;
    }
    if (B > c->Bcap) {
        free_batch_group(c);
        for (int l = 0; l < c->nlev; ++l) {
            Level& L = c->lev[l];
            TRY(dev_alloc(c, &L.I, (size_t)(B + 1) * L.plane));
            TRY(dev_alloc(c, &L.fA, (size_t)B * L.fp * L.h));
            TRY(dev_alloc(c, &L.fB, (size_t)B * L.fp * L.h));
        }
        c->S = B + 2;      // the first batch of a bracket may carry B + 1 frames (= B pairs)
        TRY(dev_alloc(c, &c->R, (size_t)c->S * c->r_slot_floats));
        const size_t fbytes = (size_t)W * H;
        for (int b = 0; b < 2; ++b) {
            TRY(dev_alloc(c, &c->d_u8[b], (size_t)(B + 1) * fbytes));
            void* hp = nullptr;
            TRY(host_alloc(c, &hp, (size_t)(B + 1) * fbytes));
            c->h_pin[b] = (uint8_t*)hp;
        }
        c->Bcap = B;
    }
    if (need_ring > c->ring_n) {
        free_ring_group(c);
        TRY(dev_alloc(c, &c->ring, (size_t)need_ring * c->ring_stride));
        TRY(dev_alloc(c, &c->d_pkey, (size_t)need_ring * c->div_nblk));
        TRY(dev_alloc(c, &c->d_psum, (size_t)need_ring * c->div_nblk));
        TRY(dev_alloc(c, &c->d_rpart, (size_t)need_ring * c->rad_gx * c->rad_gy));
        c->ring_n = need_ring;
    }
    if (maxPairs > c->pairsCap) {
        free_pairs_group(c);
        const size_t m = (size_t)maxPairs;
        TRY(dev_alloc(c, &c->d_cx, m)); TRY(dev_alloc(c, &c->d_cy, m));
        TRY(dev_alloc(c, &c->d_val, m)); TRY(dev_alloc(c, &c->d_mm, m));
        TRY(dev_alloc(c, &c->d_cut, m));
        TRY(dev_alloc(c, &c->d_centers, 2 * m)); TRY(dev_alloc(c, &c->d_scalar, m));
        TRY(dev_alloc(c, &c->d_cx_ext, m + 12)); TRY(dev_alloc(c, &c->d_cy_ext, m + 12));
        void* hr = nullptr;
        TRY(host_alloc(c, &hr, m * 48 + 256));
        c->h_res = (char*)hr;
        c->pairsCap = maxPairs;
    }
    c->B = B;
    c->maxPairs = maxPairs;
    return FFB_OK;
}

// ------------------------------------------------------------------ the per-batch pipeline
// All levels in one pass when every level is an exact 2^k decimation (W, H multiples of 8).
bool pyramid_fast_ok(int W, int H, const LevelPlan& p) {
    static const bool generic = getenv("FFB_PYR_GENERIC") && atoi(getenv("FFB_PYR_GENERIC")) != 0;
    if (generic || W % 8 || H % 8 || W < 64 || H < 64) return false;
    static const int want[FFB_MAX_LEVELS] = {3, 3, 9, 19};
    for (int l = 0; l < p.n; ++l)
        if (p.ksize[l] != want[p.k[l]] || p.w[l] != (W >> p.k[l]) || p.h[l] != (H >> p.k[l])) return false;
    return true;
}

// dst / dstride / dp are indexed by level k (0 = full resolution)
int launch_pyramid_pow2(ffb_ctx* c, const uint8_t* src, size_t stride, int pitch, int W, int H, int nlev,
                        float* const* dst, const size_t* dstride, const int* dp, int nb) {
    FfbPyr2Args a;
    memset(&a, 0, sizeof(a));
    a.src = src; a.src_frame_stride = stride; a.src_pitch = pitch; a.W = W; a.H = H;
    a.nlev = nlev;
    a.aligned = (pitch % 4 == 0) && (stride % 4 == 0) && (((uintptr_t)src) % 4 == 0);
    static const int ks[FFB_MAX_LEVELS] = {3, 3, 9, 19};
    double px = 0;
    for (int k = 0; k < nlev; ++k) {
        a.dst[k] = dst[k]; a.dstride[k] = dstride[k]; a.dp[k] = dp[k];
        a.taps[k] = make_taps(ks[k], k == 0 ? 0.0 : ((double)(1 << k) - 1.0) * 0.5);
        px += (double)(W >> k) * (H >> k);
    }
    if (!c->attr_pyr2) {
        CK(c, cudaFuncSetAttribute(k_pyramid_pow2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PYR2_SMEM));
        c->attr_pyr2 = true;
    }
    dim3 grid((W + PYR2_TX - 1) / PYR2_TX, (H + PYR2_TY - 1) / PYR2_TY, nb);
    prof_begin(c, FFB_K_PYRAMID, (double)nb * ((double)W * H + 4.0 * px));
    FFB_LAUNCH(k_pyramid_pow2, grid, dim3(256), PYR2_SMEM, c->s_comp, a);
    prof_end(c);
    CKL(c);
    return FFB_OK;
}

int expand_frames(ffb_ctx* c, const uint8_t* src, size_t stride, int pitch, int nb, int first_frame) {
    const bool fast = c->pyr_fast;
    if (fast) {
        float* dst[FFB_MAX_LEVELS] = {nullptr, nullptr, nullptr, nullptr};
        size_t dstride[FFB_MAX_LEVELS] = {0, 0, 0, 0};
        int dp[FFB_MAX_LEVELS] = {0, 0, 0, 0};
        for (int l = 0; l < c->nlev; ++l) {
            Level& L = c->lev[l];
            dst[L.k] = L.I; dstride[L.k] = L.plane; dp[L.k] = L.rp;
        }
        TRY(launch_pyramid_pow2(c, src, stride, pitch, c->W, c->H, c->nlev, dst, dstride, dp, nb));
    }
    for (int l = 0; l < c->nlev; ++l) {
        Level& L = c->lev[l];
        if (!fast)
            TRY(launch_pyramid(c, src, stride, pitch, c->W, c->H, L.I, L.plane, L.rp, L.w, L.h, L.xi, L.xa, L.yi, L.ya,
                               L.taps, L.RW, L.RH, nb));
        FfbRing dst;
        dst.base = (char*)(c->R + L.r_off);
        dst.stride = c->r_slot_floats * sizeof(float);
        dst.first = first_frame % c->S;
        dst.mod = c->S;
        TRY(launch_polyexp(c, L.I, L.plane, L.rp, L.w, L.h, dst, L.plane, L.rp, nb));
    }
    return FFB_OK;
}

int launch_divmag(ffb_ctx* c, int p0, int off, int cnt, cudaStream_t stream);

int flow_pairs(ffb_ctx* c, int p0, int np) {
    c->phase_timing = false;
    c->seg_frame_px = (long long)c->W * c->H;
    static const bool fuse_up = !(getenv("FFB_FUSE_UP") && atoi(getenv("FFB_FUSE_UP")) == 0);
    // FFB_FLOW_STREAMS=n (1..4): the pairs of a batch are split in n slices whose launch chains run on
    // n streams, so that the short coarse-level launches (and the tail of every launch) of one slice
    // overlap the other slices' work.  Every buffer is indexed by pair; the slices share nothing that
    // is written.
    const int want_slices = c->flow_streams;
    bool all_fused = fuse_up;
    for (int l = 1; l < c->nlev; ++l)
        all_fused = all_fused && c->lev[l].w == 2 * c->lev[l - 1].w && c->lev[l].h == 2 * c->lev[l - 1].h;
    int nslice = want_slices < 1 ? 1 : (want_slices > 4 ? 4 : want_slices);
    if (!all_fused || np < 4 * nslice) nslice = 1;
    int soff[5];
    for (int i = 0; i <= nslice; ++i) soff[i] = (int)((long long)np * i / nslice);
    if (nslice > 1) {
        prof_phase_begin(c);
        c->phase_timing = true;
        CK(c, cudaEventRecord(c->ev_fork, c->s_comp));
        for (int i = 1; i < nslice; ++i) CK(c, cudaStreamWaitEvent(c->s_aux[i - 1], c->ev_fork, 0));
    }
    for (int l = 0; l < c->nlev; ++l) {
        Level& L = c->lev[l];
        c->cur_level = L.k;
        const size_t fstride = (size_t)L.fp * L.h;
        const bool last = l == c->nlev - 1;
        const float2* fin0 = nullptr;
        const bool fused = l > 0 && fuse_up && L.w == 2 * c->lev[l - 1].w && L.h == 2 * c->lev[l - 1].h;
        if (l > 0 && !fused) {
            Level& C = c->lev[l - 1];
            TRY(launch_upsample(c, C.fB, (size_t)C.fp * C.h, C.fp, C.w, C.h, L.fA, fstride, L.fp, L.w, L.h, L.uxi,
                                L.uxa, L.uyi, L.uya, np));
            fin0 = L.fA;
        }
        for (int it = 0; it < FFB_ITERS; ++it) {
            for (int sl = 0; sl < nslice; ++sl) {
                const int off = soff[sl], cnt = soff[sl + 1] - soff[sl];
                c->launch_stream = sl == 0 ? c->s_comp : c->s_aux[sl - 1];
                FfbRing R;
                R.base = (char*)(c->R + L.r_off);
                R.stride = c->r_slot_floats * sizeof(float);
                R.first = (p0 + off) % c->S;
                R.mod = c->S;
                FfbRing toB{(char*)(L.fB + (size_t)off * fstride), fstride * sizeof(float2), 0, 1 << 30};
                FfbRing toA{(char*)(L.fA + (size_t)off * fstride), fstride * sizeof(float2), 0, 1 << 30};
                FfbRing toRing{(char*)c->ring, c->ring_stride * sizeof(float2), (p0 + off) % c->ring_n, c->ring_n};
                int rc;
                if (it == 0) {
                    UpSrc up;
                    if (fused) {
                        Level& C = c->lev[l - 1];
                        up.stride = (size_t)C.fp * C.h;
                        up.src = C.fB + (size_t)off * up.stride; up.sp = C.fp; up.wc = C.w; up.hc = C.h;
                    }
                    rc = launch_flow_iter(c, R, L.plane, L.rp, L.w, L.h, fin0 ? fin0 + (size_t)off * fstride : nullptr, fstride,
                                          L.fp, toB, L.fp, cnt, &up);
                } else if (it == 1) {
                    rc = launch_flow_iter(c, R, L.plane, L.rp, L.w, L.h, L.fB + (size_t)off * fstride, fstride, L.fp, toA, L.fp, cnt);
                } else {
                    rc = launch_flow_iter(c, R, L.plane, L.rp, L.w, L.h, L.fA + (size_t)off * fstride, fstride, L.fp,
                                          last ? toRing : toB, L.fp, cnt);
                }
                c->launch_stream = c->s_comp;
                TRY(rc);
            }
        }
    }
    if (nslice > 1) {
        c->phase_timing = false;
        for (int sl = 0; sl < nslice; ++sl) CK(c, cudaEventRecord(c->ev_flow_end[sl], sl == 0 ? c->s_comp : c->s_aux[sl - 1]));
        prof_phase_end(c, nslice);
        // each slice reduces its own pairs as soon as its chain is done (overlaps the other slice's tail)
        for (int sl = 0; sl < nslice; ++sl)
            TRY(launch_divmag(c, p0, soff[sl], soff[sl + 1] - soff[sl], sl == 0 ? c->s_comp : c->s_aux[sl - 1]));
    }
    for (int i = 1; i < nslice; ++i) {
        CK(c, cudaEventRecord(c->ev_join[i - 1], c->s_aux[i - 1]));
        CK(c, cudaStreamWaitEvent(c->s_comp, c->ev_join[i - 1], 0));
    }
    c->sliced_divmag = nslice > 1;
    c->cur_level = -1;
    c->seg_frame_px = 0;
    return FFB_OK;
}

// divergence / magnitude partials of `cnt` pairs starting at pair p0 + off of the bracket, on `stream`
int launch_divmag(ffb_ctx* c, int p0, int off, int cnt, cudaStream_t stream) {
    FfbDivArgs d;
    d.flow = FfbRing{(char*)c->ring, c->ring_stride * sizeof(float2), (p0 + off) % c->ring_n, c->ring_n};
    d.fp = c->fp0; d.w = c->W; d.h = c->H; d.rows_per_block = c->div_rpb;
    d.pkey = c->d_pkey + (size_t)off * c->div_nblk;
    d.psum = c->d_psum + (size_t)off * c->div_nblk;
    c->aux_stream = stream;
    prof_begin(c, FFB_K_DIVMAG, (double)cnt * 8.0 * c->W * c->H);
    FFB_LAUNCH(k_divmag, dim3(c->div_gx, c->div_gy, cnt), dim3(256), 0, stream, d);
    prof_end(c);
    c->aux_stream = c->s_comp;
    CKL(c);
    return FFB_OK;
}

int launch_phase1_finish(ffb_ctx* c, int p0, int np) {
    FfbP1Args f;
    f.flow = FfbRing{(char*)c->ring, c->ring_stride * sizeof(float2), p0 % c->ring_n, c->ring_n};
    f.fp = c->fp0; f.w = c->W; f.h = c->H;
    f.pkey = c->d_pkey; f.psum = c->d_psum; f.nblk = c->div_nblk;
    f.pov = c->pov; f.cut_threshold = c->thr; f.out0 = p0;
    f.cx = c->d_cx; f.cy = c->d_cy; f.val = c->d_val; f.mean_mag = c->d_mm; f.cut = c->d_cut;
    prof_begin(c, FFB_K_SMALL, 0);
    FFB_LAUNCH(k_phase1_finish, dim3(np), dim3(32), 0, c->s_comp, f);
    prof_end(c);
    CKL(c);
    return FFB_OK;
}

int phase1_reduce(ffb_ctx* c, int p0, int np) {
    TRY(launch_divmag(c, p0, 0, np, c->s_comp));
    return launch_phase1_finish(c, p0, np);
}

// radial pass for bracket pairs [j0, j1); n = number of pairs known so far (window truncation).
// smooth = false: d_centers already holds the centres of these pairs (ffb_bracket_radial)
int radial_range(ffb_ctx* c, int j0, int j1, int n, bool smooth = true) {
    while (j0 < j1) {
        const int cnt = (j1 - j0 < c->ring_n) ? j1 - j0 : c->ring_n;
        if (smooth) {
            prof_begin(c, FFB_K_SMALL, 0);
            FFB_LAUNCH(k_smooth_centers, dim3((cnt + 127) / 128), dim3(128), 0, c->s_comp, c->d_cx, c->d_cy, n, j0,
                       j0 + cnt, 0, c->d_centers);
            prof_end(c);
            CKL(c);
        }
        FfbRadArgs r;
        r.flow = FfbRing{(char*)c->ring, c->ring_stride * sizeof(float2), j0 % c->ring_n, c->ring_n};
        r.fp = c->fp0; r.w = c->W; r.h = c->H; r.rows_per_block = c->rad_rpb;
        r.centers = c->d_centers; r.cut = c->d_cut; r.out0 = j0; r.pov = c->pov; r.partial = c->d_rpart;
        prof_begin(c, FFB_K_RADIAL, (double)cnt * 8.0 * c->W * c->H);
        FFB_LAUNCH(k_radial, dim3(c->rad_gx, c->rad_gy, cnt), dim3(256), 0, c->s_comp, r);
        prof_end(c);
        CKL(c);
        prof_begin(c, FFB_K_SMALL, 0);
        FFB_LAUNCH(k_radial_finish, dim3(cnt), dim3(32), 0, c->s_comp, (const double*)c->d_rpart,
                   c->rad_gx * c->rad_gy, c->W, c->H, j0, c->d_scalar);
        prof_end(c);
        CKL(c);
        j0 += cnt;
    }
    return FFB_OK;
}

// Pageable frames reach the pinned staging buffer through host memcpy: one thread moves ~7 GB/s, less than the
// GPU consumes at 1080p (a 257-frame bracket is 533 MB every 36 ms), so large copies are cut into row ranges and
// moved by a few threads (FFB_COPY_THREADS, default 4; they inherit the caller's CPU affinity, i.e. the NUMA
// node of the GPU when the process was bound to it).  n frames of `rows` rows of `row` bytes each.
void copy_frames(uint8_t* dst, size_t dst_frame, const uint8_t* src, size_t src_frame, size_t src_pitch, int n, int rows,
                 size_t row, int max_threads) {
    const long long total_rows = (long long)n * rows;
    auto work = [=](long long r0, long long r1) {
        while (r0 < r1) {
            const int f = (int)(r0 / rows), a = (int)(r0 % rows);
            const int b = (int)((r1 - (long long)f * rows) < rows ? (r1 - (long long)f * rows) : rows);
            const uint8_t* s0 = src + (size_t)f * src_frame + (size_t)a * src_pitch;
            uint8_t* d0 = dst + (size_t)f * dst_frame + (size_t)a * row;
            if (src_pitch == row) memcpy(d0, s0, (size_t)(b - a) * row);
            else for (int y = 0; y < b - a; ++y) memcpy(d0 + (size_t)y * row, s0 + (size_t)y * src_pitch, row);
            r0 += b - a;
        }
    };
    int t = (int)(((size_t)total_rows * row) >> 21);          // at least 2 MB per thread
    if (t > max_threads) t = max_threads;
    if (t <= 1) { work(0, total_rows); return; }
    std::vector<std::thread> pool;
    for (int i = 1; i < t; ++i) pool.emplace_back(work, total_rows * i / t, total_rows * (i + 1) / t);
    work(0, total_rows / t);
    for (std::thread& th : pool) th.join();
}

enum PtrKind { PTR_PAGEABLE, PTR_PINNED, PTR_DEVICE };
PtrKind classify(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return PTR_PAGEABLE;
    }
    if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) return PTR_DEVICE;
    if (at.type == cudaMemoryTypeHost) return PTR_PINNED;
    return PTR_PAGEABLE;
}

int process_batch(ffb_ctx* c, const uint8_t* frames, int nb, size_t pitch, size_t stride, PtrKind kind) {
    const int b = c->batch_no & 1;
    const size_t fbytes = (size_t)c->W * c->H;
    const uint8_t* src = frames;
    size_t s_stride = stride;
    int s_pitch = (int)pitch;
    if (kind != PTR_DEVICE) {
        // staging buffer b was last read by the expansion kernels two batches ago
        CK(c, cudaStreamWaitEvent(c->s_copy, c->ev_expand[b], 0));
        if (kind == PTR_PAGEABLE) {
            CK(c, cudaEventSynchronize(c->ev_h2d[b]));   // previous DMA out of h_pin[b] finished
            copy_frames(c->h_pin[b], fbytes, frames, stride, pitch, nb, c->H, (size_t)c->W, c->copy_threads);
            CK(c, cudaMemcpyAsync(c->d_u8[b], c->h_pin[b], (size_t)nb * fbytes, cudaMemcpyHostToDevice, c->s_copy));
        } else if (pitch == (size_t)c->W && stride == fbytes) {
            CK(c, cudaMemcpyAsync(c->d_u8[b], frames, (size_t)nb * fbytes, cudaMemcpyHostToDevice, c->s_copy));
        } else {
            for (int f = 0; f < nb; ++f)
                CK(c, cudaMemcpy2DAsync(c->d_u8[b] + (size_t)f * fbytes, c->W, frames + (size_t)f * stride, pitch, c->W,
                                        c->H, cudaMemcpyHostToDevice, c->s_copy));
        }
        CK(c, cudaEventRecord(c->ev_h2d[b], c->s_copy));
        CK(c, cudaStreamWaitEvent(c->s_comp, c->ev_h2d[b], 0));
        src = c->d_u8[b];
        s_stride = fbytes;
        s_pitch = c->W;
    }
    const int g0 = c->frames_seen;
    TRY(expand_frames(c, src, s_stride, s_pitch, nb, g0));
    CK(c, cudaEventRecord(c->ev_expand[b], c->s_comp));
    const int np = nb - (g0 == 0 ? 1 : 0);
    if (np > 0) {
        const int p0 = c->pairs_done;
        if (p0 + np > c->maxPairs) return fail(c, FFB_E_INVALID, "bracket exceeds max_bracket_pairs=%d", c->maxPairs);
        if (c->deferred && p0 + np > c->ring_n)
            return fail(c, FFB_E_INVALID, "shard exceeds the %d pairs announced to ffb_bracket_begin_shard", c->ring_n);
        TRY(flow_pairs(c, p0, np));
        if (c->sliced_divmag) TRY(launch_phase1_finish(c, p0, np));
        else TRY(phase1_reduce(c, p0, np));
        c->pairs_done += np;
        const int r1 = c->pairs_done - 6;
        if (!c->deferred && r1 > c->radial_done) {
            TRY(radial_range(c, c->radial_done, r1, c->pairs_done));
            c->radial_done = r1;
        }
    }
    c->frames_seen += nb;
    c->batch_no++;
    return FFB_OK;
}

int ensure_config(ffb_ctx* c, int w, int h) {
    if (c->W == w && c->H == h && c->B >= 2) return FFB_OK;
    return configure(c, w, h, 2, 64);
}

}  // namespace

namespace {
struct Scratch {   // frees everything it allocated when the hook returns
    std::vector<void*> ptrs;
    ~Scratch() { for (void* p : ptrs) cudaFree(p); }
    template <class T> int alloc(ffb_ctx* c, T** p, size_t n) {
        TRY(dev_alloc(c, p, n));
        ptrs.push_back(*p);
        return FFB_OK;
    }
    template <class T> int upload(ffb_ctx* c, T** p, const T* h, size_t n) {
        TRY(alloc(c, p, n));
        CK(c, cudaMemcpy(*p, h, n * sizeof(T), cudaMemcpyHostToDevice));
        return FFB_OK;
    }
};
}  // namespace

// ------------------------------------------------------------------ frame pre-processing (row N2)
constexpr int PRE_CHUNK = 8;      // colour frames per staging chunk
constexpr int PRE_OUT = 256;

// cv::resize INTER_LINEAR uint8 tables: [x0 | x1 | a0 | a1] (x: weights reset where the footprint
// leaves the image; y: only the indices are clipped) -- see oracle/preproc_np.py.
std::vector<int> make_u8_resize_table(int dst_n, int src_n, bool reset_at_borders) {
    std::vector<int> t(4 * (size_t)dst_n);
    const double scale = (double)src_n / dst_n;
    for (int d = 0; d < dst_n; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s0 = (int)floorf(f);
        f -= (float)s0;
        if (reset_at_borders) {
            if (s0 < 0) { s0 = 0; f = 0.f; }
            if (s0 >= src_n - 1) { s0 = src_n - 1; f = 0.f; }
        }
        const int a1 = (int)nearbyintf(f * 2048.f);
        const int a0 = (int)nearbyintf((1.f - f) * 2048.f);
        auto clip = [&](int v) { return v < 0 ? 0 : (v > src_n - 1 ? src_n - 1 : v); };
        t[d] = clip(s0);
        t[dst_n + d] = clip(s0 + 1);
        t[2 * dst_n + d] = a0;
        t[3 * dst_n + d] = a1;
    }
    return t;
}

void free_preprocess(ffb_ctx* c) {
    dev_free(c->pre_xt); dev_free(c->pre_yt);
    for (int b = 0; b < 2; ++b) {
        dev_free(c->d_color[b]);
        if (c->h_color[b]) cudaFreeHost(c->h_color[b]);
        c->h_color[b] = nullptr;
    }
    c->pre = ffb_ctx::PrePlan();
}

int launch_preprocess(ffb_ctx* c, const uint8_t* src, size_t stride, int pitch, uint8_t* dst, const ffb_ctx::PrePlan& p,
                      const int* xt, const int* yt, int n) {
    FfbPreArgs a;
    a.src = src; a.src_frame_stride = stride; a.src_pitch = pitch;
    a.dst = dst; a.dst_frame_stride = (size_t)p.OW * p.OH; a.dst_pitch = p.OW;
    a.xt = xt; a.yt = yt; a.TW = p.TW; a.TH = p.TH; a.x_off = p.x0; a.y_off = p.y0; a.OW = p.OW; a.OH = p.OH;
    // algorithmic bytes: the 2x2 footprint of every output pixel (3 bytes each) + the gray output
    prof_begin(c, FFB_K_PREPROC, (double)n * ((double)p.OW * p.OH * 13.0));
    FFB_LAUNCH(k_preprocess, dim3((p.OW + 31) / 32, (p.OH + 7) / 8, n), dim3(256), 0, c->s_comp, a);
    prof_end(c);
    CKL(c);
    return FFB_OK;
}

// the reference's two modes (F:1057, F:1076-1079) as plans
ffb_ctx::PrePlan reference_plan(int W, int H, int vr) {
    ffb_ctx::PrePlan p;
    p.W = W; p.H = H; p.OW = p.OH = PRE_OUT;
    p.TW = p.TH = vr ? 2 * PRE_OUT : PRE_OUT;
    p.x0 = 0; p.y0 = vr ? PRE_OUT : 0;
    return p;
}
bool plan_ok(const ffb_ctx::PrePlan& p) {
    return p.W >= 2 && p.H >= 2 && p.TW >= 1 && p.TH >= 1 && p.OW >= 1 && p.OH >= 1 && p.x0 >= 0 && p.y0 >= 0 &&
           p.x0 + p.OW <= p.TW && p.y0 + p.OH <= p.TH;
}

// ======================================================================================
// extern "C" ABI
// ======================================================================================
extern "C" {

int ffb_version(void) { return FFB_VERSION; }

const char* ffb_kernel_name(int kid) { return kid >= 0 && kid < FFB_K_COUNT ? kKernelNames[kid] : "?"; }

int ffb_device_count(int* n) {
    if (!n) return FFB_E_INVALID;
    int k = 0;
    if (cudaGetDeviceCount(&k) != cudaSuccess) { cudaGetLastError(); k = 0; }
    *n = k;
    return FFB_OK;
}

int ffb_device_pci_bus_id(int device, char* buf, int buf_len) {
    if (!buf || buf_len < 16) return FFB_E_INVALID;
    if (cudaDeviceGetPCIBusId(buf, buf_len, device) != cudaSuccess) { cudaGetLastError(); buf[0] = 0; return FFB_E_NODEVICE; }
    return FFB_OK;
}

int ffb_device_name(int device, char* buf, int buf_len) {
    if (!buf || buf_len < 2) return FFB_E_INVALID;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { cudaGetLastError(); buf[0] = 0; return FFB_E_NODEVICE; }
    snprintf(buf, (size_t)buf_len, "%s", prop.name);
    return FFB_OK;
}

const char* ffb_last_error(const ffb_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int ffb_create(int device, ffb_ctx** out) {
    if (!out) return FFB_E_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(nullptr, FFB_E_NODEVICE, "no CUDA device available (this library has no CPU fallback)");
    }
    if (device < 0 || device >= n) return fail(nullptr, FFB_E_INVALID, "device %d out of range (0..%d)", device, n - 1);
    ffb_ctx* c = new ffb_ctx();
    c->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->s_comp, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->s_copy, cudaStreamNonBlocking) != cudaSuccess) {
        const int rc = fail(nullptr, FFB_E_CUDA, "stream creation failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete c;
        return rc;
    }
    cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_chain, cudaEventDisableTiming);
    cudaStreamCreateWithFlags(&c->s_time, cudaStreamNonBlocking);
    for (int i = 0; i < 4; ++i) cudaEventCreateWithFlags(&c->ev_flow_end[i], cudaEventDisableTiming);
    for (int i = 0; i < 3; ++i) {
        cudaStreamCreateWithFlags(&c->s_aux[i], cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming);
    }
    c->launch_stream = c->s_comp;
    c->aux_stream = c->s_comp;
    if (const char* e = getenv("FFB_FLOW_STREAMS")) c->flow_streams = atoi(e);
    if (const char* e = getenv("FFB_COPY_THREADS")) { const int v = atoi(e); if (v >= 1 && v <= 64) c->copy_threads = v; }
    for (int b = 0; b < 2; ++b) {
        cudaEventCreateWithFlags(&c->ev_h2d[b], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->ev_expand[b], cudaEventDisableTiming);
        cudaEventRecord(c->ev_h2d[b], c->s_copy);
        cudaEventRecord(c->ev_expand[b], c->s_comp);
        cudaEventCreateWithFlags(&c->ev_ch2d[b], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->ev_pre[b], cudaEventDisableTiming);
        cudaEventRecord(c->ev_ch2d[b], c->s_copy);
        cudaEventRecord(c->ev_pre[b], c->s_comp);
    }
    c->poly = make_poly_consts();
    *out = c;
    return FFB_OK;
}

void ffb_destroy(ffb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->s_comp);
    cudaStreamSynchronize(c->s_copy);
    prof_collect(c);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : c->timers) if (e) cudaEventDestroy(e);
    free_geometry(c);
    free_preprocess(c);
    for (int b = 0; b < 2; ++b) {
        cudaEventDestroy(c->ev_h2d[b]); cudaEventDestroy(c->ev_expand[b]);
        cudaEventDestroy(c->ev_ch2d[b]); cudaEventDestroy(c->ev_pre[b]);
    }
    cudaEventDestroy(c->ev_fork);
    cudaEventDestroy(c->ev_chain);
    for (int i = 0; i < 4; ++i) cudaEventDestroy(c->ev_flow_end[i]);
    cudaStreamDestroy(c->s_time);
    for (int i = 0; i < 3; ++i) { cudaEventDestroy(c->ev_join[i]); cudaStreamDestroy(c->s_aux[i]); }
    cudaStreamDestroy(c->s_comp);
    cudaStreamDestroy(c->s_copy);
    delete c;
}

int ffb_host_alloc(void** p, size_t bytes) {
    if (!p) return FFB_E_INVALID;
    return cudaHostAlloc(p, bytes, cudaHostAllocDefault) == cudaSuccess ? FFB_OK : FFB_E_NOMEM;
}
int ffb_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? FFB_OK : FFB_E_CUDA; }

static int configure_checked(ffb_ctx* c, int w, int h, int batch_frames, int max_pairs, int ring_pairs) {
    if (!c) return FFB_E_INVALID;
    CK(c, cudaSetDevice(c->device));
    const int W0 = c->W;
    const int rc = configure(c, w, h, batch_frames, max_pairs, ring_pairs);
    // an allocation that failed half-way leaves no half-built geometry behind (argument errors change nothing)
    if (rc != FFB_OK && rc != FFB_E_INVALID && !c->in_bracket) {
        const std::string msg = c->err;
        free_geometry(c);
        c->err = msg;
    }
    (void)W0;
    return rc;
}

int ffb_configure(ffb_ctx* c, int w, int h, int batch_frames, int max_pairs) {
    return configure_checked(c, w, h, batch_frames, max_pairs, 0);
}

int ffb_get_geometry(const ffb_ctx* c, int* w, int* h, int* batch_frames, int* max_pairs) {
    if (!c) return FFB_E_INVALID;
    if (w) *w = c->W;
    if (h) *h = c->H;
    if (batch_frames) *batch_frames = c->B;
    if (max_pairs) *max_pairs = c->maxPairs;
    return FFB_OK;
}

int ffb_alloc_counts(const ffb_ctx* c, int64_t* device_allocs, int64_t* host_allocs) {
    if (!c) return FFB_E_INVALID;
    if (device_allocs) *device_allocs = c->n_dev_alloc;
    if (host_allocs) *host_allocs = c->n_host_alloc;
    return FFB_OK;
}

int ffb_bracket_begin(ffb_ctx* c, int pov, double thr) {
    if (!c || c->W == 0) return fail(c, FFB_E_INVALID, "ffb_bracket_begin before ffb_configure");
    if (c->in_bracket) return fail(c, FFB_E_INVALID, "ffb_bracket_begin inside a bracket (finish or abort it first)");
    CK(c, cudaSetDevice(c->device));
    c->in_bracket = true;
    c->deferred = false;
    c->phase1_read = false;
    c->pend = 0;
    c->pov = pov ? 1 : 0;
    c->thr = (float)thr;
    c->frames_seen = c->pairs_done = c->radial_done = 0;
    return FFB_OK;
}

int ffb_bracket_begin_shard(ffb_ctx* c, int pov, double thr, int shard_pairs) {
    if (!c || c->W == 0) return fail(c, FFB_E_INVALID, "ffb_bracket_begin_shard before ffb_configure");
    if (c->in_bracket) return fail(c, FFB_E_INVALID, "ffb_bracket_begin_shard inside a bracket (finish or abort it first)");
    if (shard_pairs < 1) return fail(c, FFB_E_INVALID, "ffb_bracket_begin_shard: shard_pairs must be >= 1");
    // every final flow of the shard stays resident until the centres of its neighbours are known
    const int pairs = c->maxPairs > shard_pairs ? c->maxPairs : shard_pairs;
    TRY(configure_checked(c, c->W, c->H, c->B, pairs, shard_pairs));
    TRY(ffb_bracket_begin(c, pov, thr));
    c->deferred = true;
    return FFB_OK;
}

// Frames the next batch of the bracket may take.
static int batch_cap(const ffb_ctx* c, PtrKind kind) {
    // a bracket of k*B pairs has k*B + 1 frames: let its first batch take one frame more
    // so that no degenerate one-frame batch is left over
    if (kind == PTR_DEVICE || c->B < 16) return c->frames_seen == 0 ? c->B + 1 : c->B;
    // Host input: the upload of batch i+1 hides behind the kernels of batch i only while it is not longer than
    // them.  A GPU consumes 1080p frames at ~15 GB/s and receives them at 25 .. 55 GB/s (eight GPUs uploading at
    // once share the host's memory and PCIe roots), so batches may grow by a factor of two, not more: B/4 + 1, B/2,
    // B, B, ...  (with a jump from B/4 + 1 straight to B = 128 the second upload took ~10 ms against 4.4 ms of
    // kernels on an 8-GPU box: end-to-end 0.84 of device-resident; a ramp from B/8 + 1 costs 1.4 % on one GPU because
    // of its small batches -- profiles/r2_scaling.txt).  A 3000-frame bracket pays this ramp once.
    if (c->frames_seen == 0) return c->B / 4 + 1;
    return c->frames_seen <= c->B / 4 + 1 ? c->B / 2 : c->B;
}

// ffb_bracket_push_bgr gathers pre-processed gray frames in the device staging buffer until a batch is full;
// whatever is waiting there runs as a (short) batch before results are read or gray frames are pushed directly
static int flush_pending(ffb_ctx* c) {
    if (c->pend <= 0) return FFB_OK;
    const int nb = c->pend;
    c->pend = 0;
    return process_batch(c, c->d_u8[c->batch_no & 1], nb, (size_t)c->W, (size_t)c->W * c->H, PTR_DEVICE);
}

int ffb_bracket_push(ffb_ctx* c, const uint8_t* frames, int n, size_t pitch, size_t stride) {
    if (!c || !c->in_bracket) return fail(c, FFB_E_INVALID, "ffb_bracket_push outside a bracket");
    CK(c, cudaSetDevice(c->device));      // several contexts (GPUs) may be driven by one host thread
    if (c->phase1_read) return fail(c, FFB_E_INVALID, "ffb_bracket_push after ffb_bracket_phase1_finish");
    if (!frames || n < 0 || pitch < (size_t)c->W || stride < pitch * (size_t)(c->H - 1) + c->W)
        return fail(c, FFB_E_INVALID, "ffb_bracket_push: bad arguments");
    TRY(flush_pending(c));
    const PtrKind kind = classify(frames);
    for (int i = 0; i < n;) {
        const int cap = batch_cap(c, kind);
        const int nb = n - i < cap ? n - i : cap;
        TRY(process_batch(c, frames + (size_t)i * stride, nb, pitch, stride, kind));
        i += nb;
    }
    return FFB_OK;
}

int ffb_bracket_abort(ffb_ctx* c) {
    if (!c) return FFB_E_INVALID;
    if (!c->in_bracket) return FFB_OK;
    cudaSetDevice(c->device);
    // queued work reads the caller's buffers and the staging rings: let it drain (errors are already recorded)
    cudaStreamSynchronize(c->s_copy);
    cudaStreamSynchronize(c->s_comp);
    for (int i = 0; i < 3; ++i) cudaStreamSynchronize(c->s_aux[i]);
    cudaGetLastError();
    c->in_bracket = false;
    c->deferred = c->phase1_read = false;
    c->pend = 0;
    c->frames_seen = c->pairs_done = c->radial_done = 0;
    return FFB_OK;
}

int ffb_chain_after(ffb_ctx* later, ffb_ctx* earlier) {
    if (!later || !earlier || later == earlier) return FFB_E_INVALID;
    if (later->device != earlier->device) return fail(later, FFB_E_INVALID, "ffb_chain_after: contexts on different devices");
    CK(later, cudaSetDevice(later->device));
    // the kernels `later` queues from now on start after everything `earlier` has queued so far; uploads (s_copy) do not wait
    CK(later, cudaEventRecord(earlier->ev_chain, earlier->s_comp));
    CK(later, cudaStreamWaitEvent(later->s_comp, earlier->ev_chain, 0));
    return FFB_OK;
}

int ffb_sync(ffb_ctx* c) {
    if (!c) return FFB_E_INVALID;
    CK(c, cudaSetDevice(c->device));      // several contexts (GPUs) may be driven by one host thread
    CK(c, cudaStreamSynchronize(c->s_copy));
    CK(c, cudaStreamSynchronize(c->s_comp));
    return FFB_OK;
}

// D2H of the per-pair results of the open bracket (each output may be NULL) through the pinned staging
// buffer, then a wait for both streams.  with_radial: scalar and centers are valid.
static int fetch_results(ffb_ctx* c, int n, bool with_radial, double* scalar, uint8_t* cut, int32_t* cx, int32_t* cy, float* val,
                         float* mean_mag, double* centers) {
    if (n > 0) {
        char* h = c->h_res;
        size_t o = 0;
        double* h_scalar = (double*)(h + o); o += (size_t)n * 8;
        double* h_centers = (double*)(h + o); o += (size_t)n * 16;
        int* h_cx = (int*)(h + o); o += (size_t)n * 4;
        int* h_cy = (int*)(h + o); o += (size_t)n * 4;
        float* h_val = (float*)(h + o); o += (size_t)n * 4;
        float* h_mm = (float*)(h + o); o += (size_t)n * 4;
        unsigned char* h_cut = (unsigned char*)(h + o);
        if (with_radial) {
            CK(c, cudaMemcpyAsync(h_scalar, c->d_scalar, (size_t)n * 8, cudaMemcpyDeviceToHost, c->s_comp));
            CK(c, cudaMemcpyAsync(h_centers, c->d_centers, (size_t)n * 16, cudaMemcpyDeviceToHost, c->s_comp));
        }
        CK(c, cudaMemcpyAsync(h_cx, c->d_cx, (size_t)n * 4, cudaMemcpyDeviceToHost, c->s_comp));
        CK(c, cudaMemcpyAsync(h_cy, c->d_cy, (size_t)n * 4, cudaMemcpyDeviceToHost, c->s_comp));
        CK(c, cudaMemcpyAsync(h_val, c->d_val, (size_t)n * 4, cudaMemcpyDeviceToHost, c->s_comp));
        CK(c, cudaMemcpyAsync(h_mm, c->d_mm, (size_t)n * 4, cudaMemcpyDeviceToHost, c->s_comp));
        CK(c, cudaMemcpyAsync(h_cut, c->d_cut, (size_t)n, cudaMemcpyDeviceToHost, c->s_comp));
        CK(c, cudaStreamSynchronize(c->s_comp));
        if (with_radial && scalar) memcpy(scalar, h_scalar, (size_t)n * 8);
        if (with_radial && centers) memcpy(centers, h_centers, (size_t)n * 16);
        if (cx) memcpy(cx, h_cx, (size_t)n * 4);
        if (cy) memcpy(cy, h_cy, (size_t)n * 4);
        if (val) memcpy(val, h_val, (size_t)n * 4);
        if (mean_mag) memcpy(mean_mag, h_mm, (size_t)n * 4);
        if (cut) memcpy(cut, h_cut, (size_t)n);
    } else {
        CK(c, cudaStreamSynchronize(c->s_comp));
    }
    CK(c, cudaStreamSynchronize(c->s_copy));
    return FFB_OK;
}

int ffb_bracket_finish(ffb_ctx* c, int* n_pairs, double* scalar, uint8_t* cut, int32_t* cx, int32_t* cy, float* val,
                       float* mean_mag, double* centers) {
    if (!c || !c->in_bracket) return fail(c, FFB_E_INVALID, "ffb_bracket_finish outside a bracket");
    CK(c, cudaSetDevice(c->device));      // several contexts (GPUs) may be driven by one host thread
    if (c->deferred) return fail(c, FFB_E_INVALID, "shard bracket: use ffb_bracket_phase1_finish + ffb_bracket_radial");
    TRY(flush_pending(c));
    const int n = c->pairs_done;
    if (n > c->radial_done) {
        TRY(radial_range(c, c->radial_done, n, n));
        c->radial_done = n;
    }
    c->in_bracket = false;
    if (n_pairs) *n_pairs = n;
    return fetch_results(c, n, true, scalar, cut, cx, cy, val, mean_mag, centers);
}

int ffb_bracket_phase1_finish(ffb_ctx* c, int* n_pairs, int32_t* cx, int32_t* cy, float* val, float* mean_mag, uint8_t* cut) {
    if (!c || !c->in_bracket || !c->deferred)
        return fail(c, FFB_E_INVALID, "ffb_bracket_phase1_finish outside a bracket opened with ffb_bracket_begin_shard");
    CK(c, cudaSetDevice(c->device));      // several contexts (GPUs) may be driven by one host thread
    TRY(flush_pending(c));
    c->phase1_read = true;
    if (n_pairs) *n_pairs = c->pairs_done;
    return fetch_results(c, c->pairs_done, false, nullptr, cut, cx, cy, val, mean_mag, nullptr);
}

int ffb_bracket_radial(ffb_ctx* c, const int32_t* cx_ext, const int32_t* cy_ext, int n_ext, int first, double* scalar,
                       double* centers) {
    if (!c || !c->in_bracket || !c->deferred || !c->phase1_read)
        return fail(c, FFB_E_INVALID, "ffb_bracket_radial before ffb_bracket_phase1_finish");
    CK(c, cudaSetDevice(c->device));      // several contexts (GPUs) may be driven by one host thread
    const int n = c->pairs_done;
    if (n > 0 && (!cx_ext || !cy_ext || first < 0 || first > 6 || n_ext < first + n || n_ext > first + n + 6))
        return fail(c, FFB_E_INVALID, "ffb_bracket_radial: %d external centres with the shard's first pair at %d do not frame %d pairs",
                    n_ext, first, n);
    if (n > 0) {
        // the raw centres of the shard and of up to 6 neighbours on each side, as the neighbours computed them
        CK(c, cudaMemcpyAsync(c->d_cx_ext, cx_ext, (size_t)n_ext * 4, cudaMemcpyHostToDevice, c->s_comp));
        CK(c, cudaMemcpyAsync(c->d_cy_ext, cy_ext, (size_t)n_ext * 4, cudaMemcpyHostToDevice, c->s_comp));
        prof_begin(c, FFB_K_SMALL, 0);
        FFB_LAUNCH(k_smooth_centers, dim3((n + 127) / 128), dim3(128), 0, c->s_comp, c->d_cx_ext, c->d_cy_ext, n_ext, first,
                   first + n, first, c->d_centers);
        prof_end(c);
        CKL(c);
        TRY(radial_range(c, 0, n, n, false));
        c->radial_done = n;
    }
    c->in_bracket = false;
    c->deferred = c->phase1_read = false;
    return fetch_results(c, n, true, scalar, nullptr, nullptr, nullptr, nullptr, nullptr, centers);
}

int ffb_flow_ring_size(const ffb_ctx* c) { return c ? c->ring_n : 0; }

int ffb_bracket_get_flow(ffb_ctx* c, int pair, float* out) {
    if (!c || !out || c->W == 0) return FFB_E_INVALID;
    CK(c, cudaSetDevice(c->device));      // several contexts (GPUs) may be driven by one host thread
    if (pair < 0 || pair >= c->pairs_done || pair < c->pairs_done - c->ring_n)
        return fail(c, FFB_E_RANGE, "flow of pair %d is not resident (pairs %d, ring %d)", pair, c->pairs_done, c->ring_n);
    CK(c, cudaStreamSynchronize(c->s_comp));
    const float2* src = c->ring + (size_t)(pair % c->ring_n) * c->ring_stride;
    CK(c, cudaMemcpy2D(out, (size_t)c->W * 8, src, (size_t)c->fp0 * 8, (size_t)c->W * 8, c->H, cudaMemcpyDeviceToHost));
    return FFB_OK;
}

// ---- per-call functions ------------------------------------------------------------------
int ffb_farneback(ffb_ctx* c, const uint8_t* prev, const uint8_t* next, int w, int h, size_t pitch, float* flow) {
    if (!c || !prev || !next || !flow) return FFB_E_INVALID;
    CK(c, cudaSetDevice(c->device));
    TRY(ensure_config(c, w, h));
    TRY(ffb_bracket_begin(c, 0, std::numeric_limits<float>::infinity()));
    TRY(ffb_bracket_push(c, prev, 1, pitch, pitch * h));
    TRY(ffb_bracket_push(c, next, 1, pitch, pitch * h));
    int n = 0;
    TRY(ffb_bracket_finish(c, &n, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
    return ffb_bracket_get_flow(c, 0, flow);
}

static int upload_flow(ffb_ctx* c, const float* flow, int w, int h) {
    TRY(ensure_config(c, w, h));
    if (c->in_bracket) return fail(c, FFB_E_INVALID, "per-call function inside a bracket");
    CK(c, cudaMemcpy2DAsync(c->ring, (size_t)c->fp0 * 8, flow, (size_t)w * 8, (size_t)w * 8, h, cudaMemcpyHostToDevice, c->s_comp));
    c->pairs_done = 0;   // ring contents no longer belong to a bracket
    return FFB_OK;
}

static int flow_stats(ffb_ctx* c, const float* flow, int w, int h, int32_t* x, int32_t* y, float* val, float* mm) {
    if (!c || !flow) return FFB_E_INVALID;
    CK(c, cudaSetDevice(c->device));
    TRY(upload_flow(c, flow, w, h));
    const int pov = c->pov;
    const float thr = c->thr;
    c->pov = 0;
    c->thr = std::numeric_limits<float>::infinity();
    const int rc = phase1_reduce(c, 0, 1);
    c->pov = pov;
    c->thr = thr;
    TRY(rc);
    int hx = 0, hy = 0;
    float hv = 0, hm = 0;
    CK(c, cudaMemcpyAsync(&hx, c->d_cx, 4, cudaMemcpyDeviceToHost, c->s_comp));
    CK(c, cudaMemcpyAsync(&hy, c->d_cy, 4, cudaMemcpyDeviceToHost, c->s_comp));
    CK(c, cudaMemcpyAsync(&hv, c->d_val, 4, cudaMemcpyDeviceToHost, c->s_comp));
    CK(c, cudaMemcpyAsync(&hm, c->d_mm, 4, cudaMemcpyDeviceToHost, c->s_comp));
    CK(c, cudaStreamSynchronize(c->s_comp));
    if (x) *x = hx;
    if (y) *y = hy;
    if (val) *val = hv;
    if (mm) *mm = hm;
    return FFB_OK;
}

int ffb_max_divergence(ffb_ctx* c, const float* flow, int w, int h, int32_t* x, int32_t* y, float* val) {
    return flow_stats(c, flow, w, h, x, y, val, nullptr);
}
int ffb_mean_magnitude(ffb_ctx* c, const float* flow, int w, int h, float* mm) {
    return flow_stats(c, flow, w, h, nullptr, nullptr, nullptr, mm);
}

int ffb_radial_motion(ffb_ctx* c, const float* flow, int w, int h, double cx, double cy, int is_cut, int pov, double* out) {
    if (!c || !flow || !out) return FFB_E_INVALID;
    CK(c, cudaSetDevice(c->device));
    if (is_cut) { *out = 0.0; return FFB_OK; }   // F:766-767
    TRY(upload_flow(c, flow, w, h));
    const double cen[2] = {cx, cy};
    const unsigned char zero = 0;
    CK(c, cudaMemcpyAsync(c->d_centers, cen, 16, cudaMemcpyHostToDevice, c->s_comp));
    CK(c, cudaMemcpyAsync(c->d_cut, &zero, 1, cudaMemcpyHostToDevice, c->s_comp));
    FfbRadArgs r;
    r.flow = FfbRing{(char*)c->ring, c->ring_stride * sizeof(float2), 0, c->ring_n};
    r.fp = c->fp0; r.w = w; r.h = h; r.rows_per_block = c->rad_rpb;
    r.centers = c->d_centers; r.cut = c->d_cut; r.out0 = 0; r.pov = pov ? 1 : 0; r.partial = c->d_rpart;
    prof_begin(c, FFB_K_RADIAL, 8.0 * w * h);
    FFB_LAUNCH(k_radial, dim3(c->rad_gx, c->rad_gy, 1), dim3(256), 0, c->s_comp, r);
    prof_end(c);
    CKL(c);
    prof_begin(c, FFB_K_SMALL, 0);
    FFB_LAUNCH(k_radial_finish, dim3(1), dim3(32), 0, c->s_comp, (const double*)c->d_rpart, c->rad_gx * c->rad_gy, w, h, 0,
               c->d_scalar);
    prof_end(c);
    CKL(c);
    CK(c, cudaMemcpyAsync(out, c->d_scalar, 8, cudaMemcpyDeviceToHost, c->s_comp));
    CK(c, cudaStreamSynchronize(c->s_comp));
    return FFB_OK;
}

// ---- stage hooks ---------------------------------------------------------------------------
int ffb_level_plan(int W, int H, int* n, int* w, int* h, int* ksize, double* sigma) {
    if (!n || W < 1 || H < 1) return FFB_E_INVALID;
    const LevelPlan p = make_plan(W, H);
    *n = p.n;
    for (int i = 0; i < p.n; ++i) {
        if (w) w[i] = p.w[i];
        if (h) h[i] = p.h[i];
        if (ksize) ksize[i] = p.ksize[i];
        if (sigma) sigma[i] = p.sigma[i];
    }
    return FFB_OK;
}


int ffb_stage_pyramid(ffb_ctx* c, const uint8_t* img, int W, int H, size_t pitch, int level_k, float* out) {
    if (!c || !img || !out) return FFB_E_INVALID;
    CK(c, cudaSetDevice(c->device));
    const LevelPlan p = make_plan(W, H);
    int li = -1;
    for (int i = 0; i < p.n; ++i) if (p.k[i] == level_k) li = i;
    if (li < 0) return fail(c, FFB_E_INVALID, "level %d does not exist for %dx%d", level_k, W, H);
    const int w = p.w[li], h = p.h[li];
    const FfbTaps taps = make_taps(p.ksize[li], p.sigma[li]);
    std::vector<int> xi, yi;
    std::vector<float> xa, ya;
    make_linear_table(w, W, xi, xa);
    make_linear_table(h, H, yi, ya);
    int RW, RH;
    pyramid_window(xi, w, W, PYR_TW, taps.r, &RW);
    pyramid_window(yi, h, H, PYR_TH, taps.r, &RH);
    Scratch s;
    uint8_t* d_img = nullptr; float* d_out = nullptr; int *dxi = nullptr, *dyi = nullptr; float *dxa = nullptr, *dya = nullptr;
    TRY(s.alloc(c, &d_img, (size_t)W * H));
    CK(c, cudaMemcpy2D(d_img, W, img, pitch, W, H, cudaMemcpyHostToDevice));
    if (pyramid_fast_ok(W, H, p)) {   // the production path for this geometry: all levels in one pass
        float* dst[FFB_MAX_LEVELS] = {nullptr, nullptr, nullptr, nullptr};
        size_t dstride[FFB_MAX_LEVELS] = {0, 0, 0, 0};
        int dp[FFB_MAX_LEVELS] = {0, 0, 0, 0};
        for (int k = 0; k < p.n; ++k) {
            TRY(s.alloc(c, &dst[k], (size_t)(W >> k) * (H >> k)));
            dp[k] = W >> k;
        }
        TRY(launch_pyramid_pow2(c, d_img, (size_t)W * H, W, W, H, p.n, dst, dstride, dp, 1));
        CK(c, cudaStreamSynchronize(c->s_comp));
        CK(c, cudaMemcpy(out, dst[level_k], (size_t)w * h * 4, cudaMemcpyDeviceToHost));
        return FFB_OK;
    }
    TRY(s.alloc(c, &d_out, (size_t)w * h));
    TRY(s.upload(c, &dxi, xi.data(), xi.size())); TRY(s.upload(c, &dxa, xa.data(), xa.size()));
    TRY(s.upload(c, &dyi, yi.data(), yi.size())); TRY(s.upload(c, &dya, ya.data(), ya.size()));
    TRY(launch_pyramid(c, d_img, (size_t)W * H, W, W, H, d_out, (size_t)w * h, w, w, h, dxi, dxa, dyi, dya, taps, RW, RH, 1));
    CK(c, cudaStreamSynchronize(c->s_comp));
    CK(c, cudaMemcpy(out, d_out, (size_t)w * h * 4, cudaMemcpyDeviceToHost));
    return FFB_OK;
}

// Host-side conversion between the [5][h][w] planes of the hooks' interface and the device layout of
// an expansion (float4 [h][rp] = channels 0..3, then float [h][rp] = channel 4).
static void planes_to_expansion(const float* planes, int w, int h, int rp, std::vector<float>& dev) {
    const size_t plane = (size_t)rp * h;
    dev.assign(5 * plane, 0.f);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const size_t pix = (size_t)y * rp + x, src = (size_t)y * w + x;
            for (int ch = 0; ch < 4; ++ch) dev[4 * pix + ch] = planes[(size_t)ch * w * h + src];
            dev[4 * plane + pix] = planes[(size_t)4 * w * h + src];
        }
}
static void expansion_to_planes(const std::vector<float>& dev, int w, int h, int rp, float* planes) {
    const size_t plane = (size_t)rp * h;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const size_t pix = (size_t)y * rp + x, dst = (size_t)y * w + x;
            for (int ch = 0; ch < 4; ++ch) planes[(size_t)ch * w * h + dst] = dev[4 * pix + ch];
            planes[(size_t)4 * w * h + dst] = dev[4 * plane + pix];
        }
}

int ffb_stage_polyexp(ffb_ctx* c, const float* img, int w, int h, float* out) {
    if (!c || !img || !out) return FFB_E_INVALID;
    CK(c, cudaSetDevice(c->device));
    Scratch s;
    const int rp = ffb_round_up(w, 4);            // the kernel stores 16-byte vectors
    const size_t plane = (size_t)rp * h;
    float *d_in = nullptr, *d_out = nullptr;
    TRY(s.upload(c, &d_in, img, (size_t)w * h));
    TRY(s.alloc(c, &d_out, 5 * plane));
    FfbRing dst{(char*)d_out, 0, 0, 1};
    TRY(launch_polyexp(c, d_in, (size_t)w * h, w, w, h, dst, plane, rp, 1));
    CK(c, cudaStreamSynchronize(c->s_comp));
    std::vector<float> host(5 * plane);
    CK(c, cudaMemcpy(host.data(), d_out, 5 * plane * sizeof(float), cudaMemcpyDeviceToHost));
    expansion_to_planes(host, w, h, rp, out);
    return FFB_OK;
}

int ffb_stage_update_matrices(ffb_ctx* c, const float* R0, const float* R1, const float* flow, int w, int h, float* out) {
    if (!c || !R0 || !R1 || !out) return FFB_E_INVALID;
    CK(c, cudaSetDevice(c->device));
    Scratch s;
    float *d0 = nullptr, *d1 = nullptr, *dM = nullptr; float2* df = nullptr;
    TRY(s.upload(c, &d0, R0, (size_t)5 * w * h));
    TRY(s.upload(c, &d1, R1, (size_t)5 * w * h));
    if (flow) TRY(s.upload(c, &df, (const float2*)flow, (size_t)w * h));
    TRY(s.alloc(c, &dM, (size_t)5 * w * h));
    FfbMatArgs a{d0, d1, (size_t)w * h, w, w, h, df, w, dM};
    prof_begin(c, FFB_K_SMALL, 0);
    FFB_LAUNCH(k_update_matrices, dim3((w + 31) / 32, (h + 7) / 8), dim3(256), 0, c->s_comp, a);
    prof_end(c);
    CKL(c);
    CK(c, cudaStreamSynchronize(c->s_comp));
    CK(c, cudaMemcpy(out, dM, (size_t)5 * w * h * 4, cudaMemcpyDeviceToHost));
    return FFB_OK;
}

int ffb_stage_flow_iter(ffb_ctx* c, const float* R0, const float* R1, const float* flow_in, int w, int h, float* flow_out) {
    if (!c || !R0 || !R1 || !flow_out) return FFB_E_INVALID;
    CK(c, cudaSetDevice(c->device));
    Scratch s;
    const int rp = ffb_round_up(w, 4);
    const size_t plane = (size_t)rp * h;
    float* dR = nullptr; float2 *dfi = nullptr, *dfo = nullptr;
    TRY(s.alloc(c, &dR, 10 * plane));
    for (int f = 0; f < 2; ++f) {
        std::vector<float> host;
        planes_to_expansion(f ? R1 : R0, w, h, rp, host);
        CK(c, cudaMemcpy(dR + (size_t)f * 5 * plane, host.data(), 5 * plane * sizeof(float), cudaMemcpyHostToDevice));
    }
    if (flow_in) {
        TRY(s.alloc(c, &dfi, plane));
        CK(c, cudaMemcpy2D(dfi, (size_t)rp * 8, flow_in, (size_t)w * 8, (size_t)w * 8, h, cudaMemcpyHostToDevice));
    }
    TRY(s.alloc(c, &dfo, plane));
    FfbRing R{(char*)dR, 5 * plane * sizeof(float), 0, 2};
    FfbRing fo{(char*)dfo, 0, 0, 1};
    TRY(launch_flow_iter(c, R, plane, rp, w, h, dfi, 0, rp, fo, rp, 1));
    CK(c, cudaStreamSynchronize(c->s_comp));
    CK(c, cudaMemcpy2D(flow_out, (size_t)w * 8, dfo, (size_t)rp * 8, (size_t)w * 8, h, cudaMemcpyDeviceToHost));
    return FFB_OK;
}

int ffb_stage_upsample_flow(ffb_ctx* c, const float* flow_c, int wc, int hc, int w, int h, float* out) {
    if (!c || !flow_c || !out) return FFB_E_INVALID;
    CK(c, cudaSetDevice(c->device));
    std::vector<int> xi, yi;
    std::vector<float> xa, ya;
    make_linear_table(w, wc, xi, xa);
    make_linear_table(h, hc, yi, ya);
    Scratch s;
    float2 *dsrc = nullptr, *ddst = nullptr; int *dxi = nullptr, *dyi = nullptr; float *dxa = nullptr, *dya = nullptr;
    TRY(s.upload(c, &dsrc, (const float2*)flow_c, (size_t)wc * hc));
    TRY(s.alloc(c, &ddst, (size_t)w * h));
    TRY(s.upload(c, &dxi, xi.data(), xi.size())); TRY(s.upload(c, &dxa, xa.data(), xa.size()));
    TRY(s.upload(c, &dyi, yi.data(), yi.size())); TRY(s.upload(c, &dya, ya.data(), ya.size()));
    TRY(launch_upsample(c, dsrc, 0, wc, wc, hc, ddst, 0, w, w, h, dxi, dxa, dyi, dya, 1));
    CK(c, cudaStreamSynchronize(c->s_comp));
    CK(c, cudaMemcpy(out, ddst, (size_t)w * h * 8, cudaMemcpyDeviceToHost));
    return FFB_OK;
}

// ---- frame pre-processing (row N2) --------------------------------------------------------
static int preprocess_configure_plan(ffb_ctx* c, const ffb_ctx::PrePlan& p) {
    if (!c || !plan_ok(p)) return fail(c, FFB_E_INVALID, "ffb_preprocess_configure: bad geometry");
    CK(c, cudaSetDevice(c->device));
    if (c->in_bracket) return fail(c, FFB_E_INVALID, "ffb_preprocess_configure inside a bracket");
    if (c->pre == p) return FFB_OK;
    CK(c, cudaStreamSynchronize(c->s_comp));
    CK(c, cudaStreamSynchronize(c->s_copy));
    free_preprocess(c);
    TRY(upload_vec(c, &c->pre_xt, make_u8_resize_table(p.TW, p.W, true)));
    TRY(upload_vec(c, &c->pre_yt, make_u8_resize_table(p.TH, p.H, false)));
    const size_t cbytes = (size_t)PRE_CHUNK * p.W * p.H * 3;
    for (int b = 0; b < 2; ++b) {
        TRY(dev_alloc(c, &c->d_color[b], cbytes));
        void* hp = nullptr;
        TRY(host_alloc(c, &hp, cbytes));
        c->h_color[b] = (uint8_t*)hp;
    }
    c->pre = p;
    return FFB_OK;
}
int ffb_preprocess_configure(ffb_ctx* c, int W, int H, int vr) {
    return preprocess_configure_plan(c, reference_plan(W, H, vr));
}
int ffb_preprocess_configure_window(ffb_ctx* c, int W, int H, int target_w, int target_h, int win_x, int win_y, int win_w,
                                    int win_h) {
    ffb_ctx::PrePlan p;
    p.W = W; p.H = H; p.TW = target_w; p.TH = target_h; p.x0 = win_x; p.y0 = win_y; p.OW = win_w; p.OH = win_h;
    return preprocess_configure_plan(c, p);
}

int ffb_bracket_push_bgr(ffb_ctx* c, const uint8_t* bgr, int n, size_t pitch, size_t stride) {
    if (!c || !c->in_bracket) return fail(c, FFB_E_INVALID, "ffb_bracket_push_bgr outside a bracket");
    CK(c, cudaSetDevice(c->device));      // several contexts (GPUs) may be driven by one host thread
    if (c->phase1_read) return fail(c, FFB_E_INVALID, "ffb_bracket_push_bgr after ffb_bracket_phase1_finish");
    const ffb_ctx::PrePlan& pp = c->pre;
    if (pp.W == 0) return fail(c, FFB_E_INVALID, "ffb_bracket_push_bgr before ffb_preprocess_configure");
    if (c->W != pp.OW || c->H != pp.OH)
        return fail(c, FFB_E_INVALID, "pre-processed frames are %dx%d: ffb_configure(ctx, %d, %d, ...)", pp.OW, pp.OH, pp.OW, pp.OH);
    const size_t row = (size_t)pp.W * 3;
    const size_t obytes = (size_t)pp.OW * pp.OH;
    if (!bgr || n < 0 || pitch < row || stride < pitch * (size_t)(pp.H - 1) + row)
        return fail(c, FFB_E_INVALID, "ffb_bracket_push_bgr: bad arguments");
    const PtrKind kind = classify(bgr);
    const size_t fbytes = row * pp.H;
    // The gray frames gather in the device staging buffer of the next batch and run through the hot path once the
    // batch is full, however the caller cuts its pushes (a decoder hands over chunks much smaller than the batches
    // that fill the GPU at 256x256); ffb_bracket_finish runs what is left.
    for (int i = 0; i < n;) {
        const int cap = batch_cap(c, kind);
        const int b = c->batch_no & 1;
        int m = n - i;
        if (m > PRE_CHUNK) m = PRE_CHUNK;
        if (m > cap - c->pend) m = cap - c->pend;
        const uint8_t* src = bgr + (size_t)i * stride;
        uint8_t* dst = c->d_u8[b] + (size_t)c->pend * obytes;
        if (kind != PTR_DEVICE) {
            const int cb = c->color_no & 1;
            c->color_no++;
            CK(c, cudaStreamWaitEvent(c->s_copy, c->ev_pre[cb], 0));      // kernel that last read d_color[cb]
            if (kind == PTR_PAGEABLE) {
                CK(c, cudaEventSynchronize(c->ev_ch2d[cb]));              // DMA that last read h_color[cb]
                copy_frames(c->h_color[cb], fbytes, src, stride, pitch, m, pp.H, row, c->copy_threads);
                CK(c, cudaMemcpyAsync(c->d_color[cb], c->h_color[cb], (size_t)m * fbytes, cudaMemcpyHostToDevice, c->s_copy));
            } else if (pitch == row && stride == fbytes) {
                CK(c, cudaMemcpyAsync(c->d_color[cb], src, (size_t)m * fbytes, cudaMemcpyHostToDevice, c->s_copy));
            } else {
                for (int f = 0; f < m; ++f)
                    CK(c, cudaMemcpy2DAsync(c->d_color[cb] + (size_t)f * fbytes, row, src + (size_t)f * stride, pitch, row,
                                            pp.H, cudaMemcpyHostToDevice, c->s_copy));
            }
            CK(c, cudaEventRecord(c->ev_ch2d[cb], c->s_copy));
            CK(c, cudaStreamWaitEvent(c->s_comp, c->ev_ch2d[cb], 0));
            TRY(launch_preprocess(c, c->d_color[cb], fbytes, (int)row, dst, pp, c->pre_xt, c->pre_yt, m));
            CK(c, cudaEventRecord(c->ev_pre[cb], c->s_comp));
        } else {
            TRY(launch_preprocess(c, src, stride, (int)pitch, dst, pp, c->pre_xt, c->pre_yt, m));
        }
        c->pend += m;
        i += m;
        // the gray frames now sit in the device staging buffer: continue as for device input
        if (c->pend >= cap) TRY(flush_pending(c));
    }
    return FFB_OK;
}

static int stage_preprocess_plan(ffb_ctx* c, const uint8_t* bgr, size_t pitch, const ffb_ctx::PrePlan& p, uint8_t* gray) {
    if (!c || !bgr || !gray || !plan_ok(p) || pitch < (size_t)p.W * 3) return fail(c, FFB_E_INVALID, "ffb_stage_preprocess: bad arguments");
    CK(c, cudaSetDevice(c->device));
    Scratch s;
    uint8_t *d_src = nullptr, *d_dst = nullptr;
    int *dxt = nullptr, *dyt = nullptr;
    const std::vector<int> xt = make_u8_resize_table(p.TW, p.W, true), yt = make_u8_resize_table(p.TH, p.H, false);
    TRY(s.alloc(c, &d_src, (size_t)p.W * p.H * 3));
    CK(c, cudaMemcpy2D(d_src, (size_t)p.W * 3, bgr, pitch, (size_t)p.W * 3, p.H, cudaMemcpyHostToDevice));
    TRY(s.alloc(c, &d_dst, (size_t)p.OW * p.OH));
    TRY(s.upload(c, &dxt, xt.data(), xt.size()));
    TRY(s.upload(c, &dyt, yt.data(), yt.size()));
    TRY(launch_preprocess(c, d_src, (size_t)p.W * p.H * 3, p.W * 3, d_dst, p, dxt, dyt, 1));
    CK(c, cudaStreamSynchronize(c->s_comp));
    CK(c, cudaMemcpy(gray, d_dst, (size_t)p.OW * p.OH, cudaMemcpyDeviceToHost));
    return FFB_OK;
}
int ffb_stage_preprocess(ffb_ctx* c, const uint8_t* bgr, int W, int H, size_t pitch, int vr, uint8_t* gray) {
    return stage_preprocess_plan(c, bgr, pitch, reference_plan(W, H, vr), gray);
}
int ffb_stage_preprocess_window(ffb_ctx* c, const uint8_t* bgr, int W, int H, size_t pitch, int target_w, int target_h, int win_x,
                                int win_y, int win_w, int win_h, uint8_t* gray) {
    ffb_ctx::PrePlan p;
    p.W = W; p.H = H; p.TW = target_w; p.TH = target_h; p.x0 = win_x; p.y0 = win_y; p.OW = win_w; p.OH = win_h;
    return stage_preprocess_plan(c, bgr, pitch, p, gray);
}

// ---- instrumentation ---------------------------------------------------------------------
int ffb_profile(ffb_ctx* c, int enable) {
    if (!c) return FFB_E_INVALID;
    CK(c, cudaStreamSynchronize(c->s_comp));
    prof_collect(c);
    c->prof = enable != 0;
    return FFB_OK;
}
int ffb_profile_reset(ffb_ctx* c) {
    if (!c) return FFB_E_INVALID;
    CK(c, cudaStreamSynchronize(c->s_comp));
    prof_collect(c);
    for (int i = 0; i < FFB_K_COUNT; ++i) { c->k_launches[i] = 0; c->k_ms[i] = 0; c->k_bytes[i] = 0; }
    for (int i = 0; i < FFB_MAX_LEVELS; ++i) { c->it_launches[i] = 0; c->it_ms[i] = 0; c->it_bytes[i] = 0; }
    return FFB_OK;
}
int ffb_kernel_stats(ffb_ctx* c, int kid, int64_t* launches, double* ms, double* bytes) {
    if (!c || kid < 0 || kid >= FFB_K_COUNT) return FFB_E_INVALID;
    CK(c, cudaStreamSynchronize(c->s_comp));
    prof_collect(c);
    if (launches) *launches = c->k_launches[kid];
    if (ms) *ms = c->k_ms[kid];
    if (bytes) *bytes = c->k_bytes[kid];
    return FFB_OK;
}
int ffb_flow_iter_level_stats(ffb_ctx* c, int k, int64_t* launches, double* ms, double* bytes) {
    if (!c || k < 0 || k >= FFB_MAX_LEVELS) return FFB_E_INVALID;
    CK(c, cudaStreamSynchronize(c->s_comp));
    prof_collect(c);
    if (launches) *launches = c->it_launches[k];
    if (ms) *ms = c->it_ms[k];
    if (bytes) *bytes = c->it_bytes[k];
    return FFB_OK;
}
int64_t ffb_launch_count(const ffb_ctx* c) { return c ? c->launches : 0; }

int ffb_timer_mark(ffb_ctx* c, int slot) {
    if (!c || slot < 0 || slot >= 8) return FFB_E_INVALID;
    if (!c->timers[slot]) CK(c, cudaEventCreate(&c->timers[slot]));
    CK(c, cudaEventRecord(c->timers[slot], c->s_comp));
    return FFB_OK;
}
int ffb_timer_elapsed(ffb_ctx* c, int a, int b, double* ms) {
    if (!c || !ms || a < 0 || a >= 8 || b < 0 || b >= 8 || !c->timers[a] || !c->timers[b]) return FFB_E_INVALID;
    CK(c, cudaEventSynchronize(c->timers[b]));
    float f = 0.f;
    CK(c, cudaEventElapsedTime(&f, c->timers[a], c->timers[b]));
    *ms = f;
    return FFB_OK;
}

}  // extern "C"
