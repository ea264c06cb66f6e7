// Context and device-side parameter blocks of the stepping engine.
#pragma once
#include <vector>

#include "../../include/pymd_b200.h"
#include "common.cuh"
#include "gemm.cuh"

namespace pymd {

constexpr int kBlockT = 16;   // time-block length of the fused path (ring sizing uses it too)
constexpr int kFarS = 32;     // super-block length of the FFT far field (far.cu)
constexpr int kFarMinMl = 24; // shorter memories stay on the direct block-Toeplitz product
constexpr int kMidK = 8 + kFarS;  // memory-kernel taps k = 2 .. 2+kMidK-1 can meet rows of the current super-block

struct BathDev {
    int nc, ncp, ml, R;
    int has_aq, nonlocal;
    const int* inv;        // [nd] -> local index or -1
    const int* cids;       // [nc]
    const double* noise;   // [nmd][nc][ldb]
    long long noise_ld;    // doubles between noise rows; 0 = nc * ldb (contiguous)
    double* cur;           // [nmd][ldb]
    double* fhis;          // [nmd][nc][ldb] or null
    double* ring;          // [R][ncp][ldb]
    double* g;             // [ncp][ldb]   head term  Hd.p_c (- Aq.q_c)
    double* uq;            // [ncp][ldb]   Aq.q'_c
    double* tail_cur;      // [T][nc][ldb]
    double* tail_next;
};

struct StepDev {
    int nd, ldb, nbath, nmd, ncon;
    long long rec_ld;      // doubles between rows of ps / qs; 0 = nd * ldb (contiguous)
    int noise_wrap;        // noise row that stands for row nmd (0: row 0, the noise is periodic in place)
    double dt;
    double *p, *q, *ps, *qs, *etot;
    double *ph, *qn, *p1, *fpotn, *fpotc, *fbase, *h;
    int* samef;            // [ldb] 1 = use cached force fpotn, 0 = use fpotc
    const double* con;     // [ncon][nd] normalised constraint vectors
    BathDev b[PYMD_MAX_BATHS];
};

struct BathHost {
    bool set = false;
    int nc = 0, ncp = 0, ml = 1, dt_scaled = 0, has_aq = 0;
    std::vector<int> cids;
    std::vector<double> kernel;   // [ml][nc][nc]
    std::vector<double> aq;       // [nc][nc]
    // device operands owned by the ctx
    int* d_cids = nullptr;
    int* d_inv = nullptr;
    double* d_head = nullptr;     // [ncp][lda]  alpha * kernel[0]
    double* d_aq = nullptr;       // [ncp][lda]
    double* d_krev = nullptr;     // [nc][Lk]    reversed, padded kernel table (tail operand)
    double* d_kintra = nullptr;   // [kBlockT][nc][nc] dt*K[1..T] for the fused path
    double* d_ghat = nullptr;     // far.cu: transformed kernel as DMMA A fragments
    double* d_far = nullptr;      // far.cu: [kFarS][nc][ldb] far field of the current super-block
    double* d_kinf = nullptr;     // fused2: dt K[1..kBlockT] as A fragments [k][ncp/8][ncp/4][32]
    double* d_hdf = nullptr;      // fused2: alpha K[0] as A fragments [ncp/8][ncp/4][32]
    double* d_tailA = nullptr;    // the two [kBlockT][nc][ldb] tail buffers inside the workspace (not owned)
    double* d_tailB = nullptr;
    double* d_kmid = nullptr;     // [kMidK][ncp/8][ncp/4][32] A fragments of dt K[k], k = 2.. (fused mid field)
    bool far_valid = false;
    long long far_base = 0;       // hcount at the start of the super-block d_far belongs to
    // fdl.cu: frequency-domain delay line of the per-step pipeline (long memories, any nc)
    int fdl_L = 0, fdl_P = 0, fdl_head = 0;   // partition length, partitions, ring slot of the newest block spectrum
    bool fdl_valid = false, fdl_tried = false;
    long long fdl_base = 0;       // hcount at the start of the super-block d_fdlF belongs to
    double* d_fdlA = nullptr;     // [L+1][2 nc][P * 2 ncp]  real block form of Ghat_p(f), partitions in reversed order
    double* d_fdlS = nullptr;     // [L+1][P][2][ncp][ldb]   spectra of the last P history blocks
    double* d_fdlY = nullptr;     // [L+1][2][nc][ldb]
    double* d_fdlF = nullptr;     // [L][nc][ldb]            far field of the current super-block
    double* d_fdlZ = nullptr;     // 3 x [2L][ncp * ldb] complex work (FFT source and ping-pong buffers)
    long long Lk = 0;
    int padf = 0;
    int lda = 0;
    int R = 0;
    long long hcount = 0;         // rows appended since the last reset
};

}  // namespace pymd

struct pymd_ctx {
    int nd = 0, ldb = 0, nbath = 0, nmd = 0, device = 0, ncon = 0;
    double dt = 0;
    int ldn = 0;                       // padded row length of nd x nd operands
    pymd::BathHost bath[PYMD_MAX_BATHS];
    std::vector<double> dyn, con, mp_host;     // mp_host: [nd][ldn] host copy of d_mp (constraint operands)
    double* d_dyn = nullptr;           // [nd][ldn]
    double* d_mp = nullptr;            // [nd][ldn]  sum_b scatter(alpha_b kernel_b[0])
    double* d_con = nullptr;           // [ncon][nd]
    double* d_aqT = nullptr;           // [nd][ldn]  q-coupled operator scattered to nd x nd (fused path)
    double* d_con3 = nullptr;          // [3][ncon][nd]  c, D c, AqT c (fused path, constraints by linearity)
    void* ws = nullptr;                // workspace
    size_t ws_bytes = 0;
    pymd::StepDev sd{};
    bool finalized = false, fpot_valid = false, tail_valid = false, fpotc_valid = false;
    long long launches = 0;
    // optional per-launch CUDA-event timing (pymd_profile_enable)
    struct ProfRec { int cls; cudaEvent_t e0, e1; };
    bool prof = false;
    std::vector<ProfRec> prof_recs;
};

namespace pymd {
bool fdl_wanted(const pymd_ctx* c, const BathHost& H);
int fdl_build(pymd_ctx* c, int b);
int fdl_advance(pymd_ctx* c, int b, cudaStream_t st);
void fdl_free(BathHost& H);
bool far_supported(const BathHost& H);
int far_build(pymd_ctx* c, int b);
int far_launch(pymd_ctx* c, unsigned mask, cudaStream_t st);
inline void prof_begin(pymd_ctx* c, int cls, cudaStream_t st) {
    if (!c->prof) return;
    pymd_ctx::ProfRec r;
    r.cls = cls;
    cudaEventCreate(&r.e0);
    cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, st);
    c->prof_recs.push_back(r);
}
inline void prof_end(pymd_ctx* c, cudaStream_t st) {
    if (c->prof && !c->prof_recs.empty()) cudaEventRecord(c->prof_recs.back().e1, st);
}
}  // namespace pymd
