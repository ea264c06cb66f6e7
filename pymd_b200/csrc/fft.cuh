// Batched FFT along the SLOW axis of [n][cols] arrays (cols = dof x trajectory, contiguous),
// Stockham autosort, every access coalesced over columns.  A pass is either a register radix-8/4/2
// butterfly per thread or a fused radix-64/16 pass (two butterfly stages exchanged through shared
// memory), so a length-2^14 transform reads and writes HBM three times instead of five.
#pragma once
#include "common.cuh"

namespace pymd {

// device table W[t] = exp(-2 pi i t / N), t < N  (cached per (device, N); host-computed)
const double2* twiddle_table(int N);

enum FftLoad { LOAD_C = 0, LOAD_REAL_PAIRS = 1 };
enum FftStore { STORE_C = 0, STORE_REAL_ROWS = 1 };

// Power-spectrum tail (spectrum.cu): when the plan holds a radix-128 pass and at least one other pass, that pass is
// ordered last and fused with the real-FFT untangling, |X_k|^2 and the column sums; `fused` reports whether it was
// (otherwise the transform is stored as asked and the caller reduces it).
struct FftPowerTail {
    int ldb, nvalid;              // column c0g + c is trajectory (c0g + c) % ldb; counted when < nvalid
    long long c0g;
    const double2* tw2;           // W_2n table
    double* partial;              // [n + 1][ncb_total]
    int ncb_total, cb0, cb_per;   // this chunk owns columns cb0 .. cb0 + cb_per - 1 of `partial`
    bool fused;
};

struct FftPlan {
    int n;                // complex length (power of two)
    long long cols;       // columns transformed
    int sign;             // -1 numpy fft, +1 numpy ifft (unscaled)
    FftLoad load;         // LOAD_REAL_PAIRS: src is real [2n][src_ld], z[m] = x[2m] + i x[2m+1]
    FftStore store;       // STORE_REAL_ROWS: dst real [2n][dst_ld]: row 2m = scale*re, row 2m+1 = scale*im
    long long src_ld, dst_ld;
    double scale;
    // STORE_REAL_ROWS only: column c of the transform lands at (c / dst_cw) * dst_cs + c % dst_cw of
    // a destination row (a chunk of dst_cw trajectories of every component row); 0 = plain c.
    long long dst_cw = 0, dst_cs = 0;
    FftPowerTail* power = nullptr;
};

// number of passes (launches) fft_columns uses for a length-n transform
int fft_num_passes(int n);

// Runs all passes.  `src` is read by the first pass only (never written unless it aliases a
// work buffer); intermediates ping-pong between bufA and bufB (each n*cols complex);
// the result of the last pass lands in `dst` (complex [n][cols] for STORE_C).
// Returns PYMD_* status; *npasses receives the number of launches.
int fft_columns(const void* src, void* dst, double2* bufA, double2* bufB, const FftPlan& plan,
                cudaStream_t stream, int* npasses);

}  // namespace pymd
