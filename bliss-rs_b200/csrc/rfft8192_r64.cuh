// rfft8192_r64.cuh -- the same 8192-point real FFT as rfft8192.cuh (4096-point complex FFT of
// z[n] = x[2n] + i x[2n+1], then the real-input untangling), cut as 4096 = 64 x 64: TWO in-register
// radix-64 passes by 64 threads per frame instead of three radix-16 passes by 256.
//
// Why: stft8192_kernel is bound by the unified L1 / shared-memory data path (ncu: 2 200 shared-memory
// wavefronts + ~1 300 L1 wavefronts per frame of 4 500 cycles).  Two passes move each complex value through
// shared memory once (+ the mirror half once more for the untangling) instead of twice and a half.
//
//   pass 1  thread b:   v[q] = z[b + 64 q]            -> Y[b][k1], times W4096^(b k1), to smem (k1, b)
//   pass 2  thread k1:  u[b] = smem (k1, b)           -> Z[k1 + 64 k2] in registers (natural order per thread)
//   untangle: bins k = k1 + 64 k2, k2 < 32, from the thread's registers; their mirrors 4096 - k =
//             (64 - k1) + 64 (63 - k2) are the upper half (k2 >= 32) of thread 64 - k1, published to smem.
//
// smem element (row r, column c) sits at r * 65 + c: column-wise (pass 1 stores, mirror loads) and row-wise
// (pass 2 loads) accesses are both conflict-free for 64-bit elements.
// __host__ __device__ so tests/cpu_emul can run the passes thread by thread.
#pragma once
#include "fft_regs.cuh"
#include "rfft8192.cuh"

namespace bliss {
namespace r64 {

constexpr int P = 65;               // row pitch in complex elements
constexpr int BUF_CPX = 64 * P;     // 4160

// pass 1: thread b in [0,64).  tw is laid out [k1][b] = W4096^(b k1) (coalesced across the lanes).
BLISS_HD void pass1_store(int b, cpx (&v)[64], const cpx *tw /*[64][64]*/, cpx *buf) {
    fft_dif64(v);
#pragma unroll
    for (int s = 0; s < 64; s++) {
        const int k1 = bitrev(s, 6);
        cpx r = v[s];
        if (k1 != 0) r = cmul(r, tw[64 * k1 + b]);
        buf[P * k1 + b] = r;
    }
}

// pass 1 with the twiddle W4096^(b k1), k1 = 8 a + c, formed as A[a] * C[c] from the thread's sixteen factors
//   A[a] = W4096^(8 a b),  C[c] = W4096^(c b)       (twf[a][b] and twf[8 + c][b], b = column: conflict-free)
// which depend on the thread only and are staged in shared memory once per CTA: the 63 table loads per
// frame (L2 latency: the 32 KB table does not stay in what is left of L1) become 16 shared-memory loads.
// Slot s holds k1 = bitrev6(s): a = bitrev3(s & 7) runs through all eight values inside each group of eight
// slots, c = bitrev3(s >> 3) is fixed per group.
BLISS_HD void pass1_store_factored(int b, cpx (&v)[64], const cpx *twf /*[16][64]*/, cpx *buf) {
    fft_dif64(v);
    const cpx *tA = twf + b, *tC = twf + 64 * 8 + b;
#pragma unroll
    for (int g = 0; g < 8; g++) {
        const int c = bitrev(g, 3);
        const cpx C = tC[64 * c];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int a = bitrev(j, 3), k1 = 8 * a + c;
            cpx r = v[8 * g + j];
            if (k1 != 0) {
                // A[a] is re-read from shared memory where it is used: v[] still fills the register file here
                const cpx w = (a == 0) ? C : (c == 0) ? tA[64 * a] : cmul(tA[64 * a], C);
                r = cmul(r, w);
            }
            buf[P * k1 + b] = r;
        }
    }
}

// pass 2: thread k1 in [0,64); afterwards u[bitrev6(k2)] = Z[k1 + 64 k2]
BLISS_HD void pass2_regs(int k1, cpx (&u)[64], const cpx *buf) {
    const cpx *p = buf + P * k1;
#pragma unroll
    for (int b = 0; b < 64; b++) u[b] = p[b];
    fft_dif64(u);
}

// publish the upper half (k2 = 32..63) of thread t: row k2 - 32, column t
BLISS_HD void publish_upper(int t, const cpx (&u)[64], cpx *buf) {
#pragma unroll
    for (int k2 = 32; k2 < 64; k2++) buf[P * (k2 - 32) + t] = u[bitrev(k2, 6)];
}

// where thread k1 finds Z[4096 - (k1 + 64 k2)], k2 < 32:  mirror_base(k1) + P * (31 - k2)
//   k1 != 0: thread 64 - k1, its k2' = 63 - k2        -> row 31 - k2, column 64 - k1
//   k1 == 0: thread 0,       its k2' = 64 - k2 (k2>0) -> row 32 - k2, column 0   (k2 == 0: Z[0] itself, own register)
BLISS_HD int mirror_base(int k1) { return k1 == 0 ? P : 64 - k1; }

}  // namespace r64
}  // namespace bliss
