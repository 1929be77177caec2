// K3, persistent form: the whole top-k eigensolve of the small Gram in ONE cooperative kernel.
//
// Reference: ruvIII.R:140 `svd(Y0 %*% t(Y0))$u[, 1:r]`; the update (ruvIII.R:142-145) uses the first k <= 32 columns.
// eigen_topk.cu runs the same block subspace iteration as ~35 strictly dependent launches with a host poll every three
// passes; at the PRPS orders (R ~ 10^3) every one of those kernels is a few microseconds of latency, so launch gaps, the
// one-CTA phases and the host round trip ARE the solve (0.41 ms at C3, half of the 8-GPU pass).  Here one grid of
// P = ceil(R / RB) CTAs (RB = 8, 16 or 32 rows of G per CTA, P <= #SMs, all co-resident: cooperative launch) keeps its
// rows of the iterate in shared memory for the whole solve and meets the other CTAs only at grid barriers:
//
//   pass:  Z_blk = G[rows, :] Q              DMMA m8n8k4, K split over the warps of the CTA, Q read from L2 (ld.cg)
//          S_c = Z_blk' Z_blk (H_c = Q_blk' Z_blk on Rayleigh-Ritz passes)          -> global partials
//          -- barrier --  every CTA reduces a slice of the P partials in a fixed order  -- barrier --
//          [Rayleigh-Ritz pass: CTA 0: H = V diag(theta) V' (parallel cyclic Jacobi) -- barrier --
//           X_blk = Q_blk V, Z_blk = Z_blk V, residual / sign partials -- barrier -- converged?  -> U, eigenvalues, exit
//           S <- V' S V]
//          every CTA: column-scaled Cholesky S = L L' (redundantly), Q_blk = Z_blk D L^-T, rows -> global Q' -- barrier --
//   [Chebyshev continuation when the first check fails: Y <- p_d(G) X with the degree-d Chebyshev polynomial that damps
//    [0, theta_min]: d products with ONE barrier each and no orthogonalisation in between (three-term recurrence on
//    the CTA's own rows), the degree chosen on the device from the Ritz values and residuals]
//
// Convergence is decided on the device; the host never waits inside the solve.  Deterministic: fixed start block, fixed
// summation orders, no floating-point atomics (the barrier counter is the only atomic).
#include <cstddef>
#include "kernels.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace ruviii {

namespace {

constexpr int PB = 32;            // block width of the subspace iteration
constexpr int PT = 256;           // threads per CTA (8 warps)
constexpr int PS = PB + 1;        // shared-memory row stride (doubles)
constexpr int kMaxTrace = 1024;

struct PersistArgs {
    const double* G;
    int R, RB, P;
    double *Q0, *Q1;              // ping-pong iterate, R x 32 row-major
    double* Pp;                   // P x 2048: per-CTA partials [S | H]
    double* Pr;                   // P x 64:   per-CTA [residual^2 partial | largest-|x| candidate] per column
    double* Pred;                 // 2048: reduced [S | H]
    double* Vm;                   // 32 x 32 Ritz rotation (row-major, columns sorted by descending theta)
    double* theta;                // 32 Ritz values, descending
    double* U;                    // out: R x kp row-major, first kk columns
    double* evals;                // out: n_keep leading Ritz values
    double* flag_out;             // optional: 0.0 when converged, 1.0 otherwise (summed by the next allreduce)
    int kk, kp, n_keep;
    int ldg;                      // row pitch of the shared-memory copy of this CTA's rows of G (GS kernels)
    double tol;
    int max_iter, first, period, cheb, jorder;
    double skip_minp;             // an orthogonalisation whose smallest pivot was above this lets the next product go without one (0: never)
    unsigned* bar;                // [0] barrier counter, [1] exit counter
    int* state;                   // [0] converged, [1] G products, [2] Cholesky breakdown, [3] Jacobi sweeps, [4] RR count,
                                  // [5] second CholQR rounds, [6] Chebyshev products
    long long* trace;             // optional: (phase << 48 | clock) marks of CTA 0
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// Grid barrier: one monotone counter.  Thread 0 arrives with a release (which orders every earlier write of the CTA, made
// visible to it by the CTA barrier, before the increment), polls with relaxed loads and closes with one acquire fence.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned& target, unsigned P) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += P;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        unsigned v;
        do {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        } while (v < target);
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
}

struct Smem {
    double Z[PB][PS];        // this CTA's rows of Z = G Q (rows >= RB unused)
    double Q[PB][PS];        // this CTA's rows of the current orthonormal block
    double X[PB][PS];        // Ritz vectors (rows) / previous Chebyshev iterate
    double A[PB][PS];        // Cholesky / Jacobi work
    double L[PB][PS];        // Cholesky factor (diagonal holds 1 / L[j][j]) / Jacobi vectors
    double S[PB][PS];        // reduced Z'Z
    double V[PB][PS];        // Ritz rotation
    double part[2304];       // K-split partials of the product (64 rows x stride 34) / reduction scratch
    double dscale[PB];
    double th[PB];
    double cs[PB / 2][2];
    int pr[PB / 2][2];
    double red[PB];
    double sgn[PB];
    int flag, flag_big, bad;
    double minp;
};

// ---- Z_blk = G[r0 : r0 + RB, :] Q -------------------------------------------------------------------------------------
// RT row tiles of 8 rows; the 8 warps are RT x KS with KS = 8 / RT splits of the contraction; partials meet in shared memory
// and are summed in a fixed order.  shift != 0: Z_blk = G Q - shift Q_own is formed later by the caller (Chebyshev).
template <int RT, bool GS>
__device__ __forceinline__ void product_phase(const PersistArgs& a, const double* __restrict__ Q, int r0, Smem& sm,
                                              const double* __restrict__ gs) {
    constexpr int KS = 8 / RT;
    constexpr int RB = 8 * RT;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, kq = lane & 3;
    const int rt = warp / KS, ks = warp % KS;
    const int R = a.R;
    const int row = r0 + rt * 8 + g;
    const bool row_ok = row < R;
    const int kchunk = ((R + KS - 1) / KS + 3) & ~3;
    const int k0 = ks * kchunk, k1 = min(R, k0 + kchunk);
    const double* grow = GS ? gs + (rt * 8 + g) * a.ldg : a.G + (int64_t)min(row, R - 1) * R;
    double acc[4][2];
#pragma unroll
    for (int t = 0; t < 4; ++t) acc[t][0] = acc[t][1] = 0.0;
    // Every CTA multiplies by the same Q at the same time: walking the contraction from a CTA-dependent offset keeps the
    // ~100 CTAs off the same L2 lines (the summation order is still fixed per CTA).
    const int nsteps = k1 > k0 ? (k1 - k0 + 3) / 4 : 0;
    int st = nsteps > 0 ? (int)((blockIdx.x * 7u) % (unsigned)nsteps) : 0;
#pragma unroll 5
    for (int i = 0; i < nsteps; ++i) {
        const int k = k0 + 4 * st;
        st = st + 1 == nsteps ? 0 : st + 1;
        const int kk = k + kq;
        const bool ok = kk < k1;
        const double av = (ok && row_ok) ? (GS ? grow[kk] : __ldg(grow + kk)) : 0.0;
        const double* qrow = Q + (int64_t)min(kk, R - 1) * PB + g;
        double b[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) b[t] = ok ? __ldcg(qrow + t * 8) : 0.0;
#pragma unroll
        for (int t = 0; t < 4; ++t) dmma884(acc[t][0], acc[t][1], av, b[t]);
    }
    constexpr int PPS = PB + 2;      // stride of the partial rows: the 8 row groups of a warp land in different banks
    double* out = sm.part + ((ks * RB) + rt * 8 + g) * PPS + kq * 2;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        out[t * 8] = acc[t][0];
        out[t * 8 + 1] = acc[t][1];
    }
    __syncthreads();
    for (int e = tid; e < RB * PB; e += PT) {
        const int r = e / PB, c = e - r * PB;
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < KS; ++q) s += sm.part[(q * RB + r) * PPS + c];
        sm.Z[r][c] = (r0 + r < R) ? s : 0.0;
    }
    __syncthreads();
}

// ---- partials of this CTA: S_c = Z_blk' Z_blk, optionally H_c = Q_blk' Z_blk ------------------------------------------
__device__ __forceinline__ void partials_phase(const PersistArgs& a, const double (*Zs)[PS], bool with_h, Smem& sm) {
    const int tid = threadIdx.x;
    double* dst = a.Pp + (int64_t)blockIdx.x * 2048;
    const int RB = a.RB;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int e = tid + u * PT, i = e / PB, j = e - i * PB;
        double s = 0.0, h = 0.0;
        for (int r = 0; r < RB; ++r) {
            const double zj = Zs[r][j];
            s = fma(Zs[r][i], zj, s);
            if (with_h) h = fma(sm.Q[r][i], zj, h);
        }
        __stcg(dst + e, s);
        if (with_h) __stcg(dst + 1024 + e, h);
    }
}

// ---- every CTA reduces a contiguous slice of the `len` partial elements over the P CTAs, in a fixed order ------------
__device__ __forceinline__ void reduce_phase(const PersistArgs& a, int len, Smem& sm) {
    const int tid = threadIdx.x, P = a.P;
    const int E = (len + P - 1) / P;                       // elements per CTA
    int epad = 1;
    while (epad < E && epad < PT) epad <<= 1;
    const int ng = PT / epad;                              // groups of partials summed side by side
    const int el = tid % epad, grp = tid / epad;
    const int e_begin = blockIdx.x * E, e_end = min(len, e_begin + E);
    for (int e0 = e_begin; e0 < e_end; e0 += epad) {
        const int e = e0 + el;
        double s = 0.0;
        if (e < e_end) {
#pragma unroll 4
            for (int p = grp; p < P; p += ng) s += __ldcg(a.Pp + (int64_t)p * 2048 + e);
        }
        sm.part[grp * epad + el] = s;
        __syncthreads();
        if (grp == 0 && e < e_end) {
            double t = 0.0;
            for (int q = 0; q < ng; ++q) t += sm.part[q * epad + el];
            __stcg(a.Pred + e, t);
        }
        __syncthreads();
    }
}

// 1 / sqrt(x) for a normal, positive x: the hardware seed and one cubic step (error^3), no range branch.
__device__ __forceinline__ double rsqrt_fast(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y * y, 1.0);
    return fma(fma(e, 0.375, 0.5), y * e, y);
}

// ---- column-scaled Cholesky of sm.S (B x B): L with 1 / L[j][j] on the diagonal, sm.dscale, sm.minp, sm.bad ------------
// One warp, lane i holds row i of the scaled matrix in registers; right-looking, column j of L travels by shuffles, so the
// 32 dependent steps cost one rsqrt + two shuffles + one FMA of latency each and no CTA barrier (a CTA-wide version with
// one barrier per column measured 35 k cycles per factorisation, five times the product with G).
__device__ __forceinline__ void cholesky_phase(Smem& sm) {
    const int tid = threadIdx.x;
    if (tid < PB) {
        const unsigned full = 0xffffffffu;
        const int i = tid;
        const double sii = sm.S[i][i];
        const double di = sii > 0.0 ? rsqrt(sii) : 1.0;
        sm.dscale[i] = di;
        __syncwarp();
        double a[PB];
#pragma unroll
        for (int c = 0; c < PB; ++c) a[c] = 0.5 * (sm.S[i][c] + sm.S[c][i]) * di * sm.dscale[c];
        int bad = 0;
        double minp = 1.0;
        // column j of L goes through shared memory (one store per lane, then broadcast 16-byte loads): per trailing column
        // one LDS.128 per two columns instead of two SHFL per column -- the step is bound by the issue rate of the
        // shuffle / shared-memory queue, not by its ~130-cycle dependency chain
        double* col = sm.part;                                // 16-byte aligned scratch, free between reduce and product
#pragma unroll
        for (int j = 0; j < PB; ++j) {
            const double pj = __shfl_sync(full, a[j], j);
            double rd = rsqrt_fast(pj);
            if (!(pj > 1e-30)) { bad = j + 1; rd = 1.0; }
            minp = fmin(minp, pj);
            const double lij = a[j] * rd;                    // L[i][j] for i >= j
            sm.L[i][j] = i > j ? lij : (i == j ? rd : 0.0);  // diagonal: 1 / L[j][j]
            if (j + 1 < PB) {
                col[(j & 1) * PB + i] = lij;
                __syncwarp();
                const double* cj = col + (j & 1) * PB;
                if (((j + 1) & 1) != 0) a[j + 1] = fma(-lij, cj[j + 1], a[j + 1]);      // odd first column, then aligned pairs
#pragma unroll
                for (int c = (j + 2) & ~1; c < PB; c += 2) {
                    const double2 l2 = *reinterpret_cast<const double2*>(cj + c);
                    a[c] = fma(-lij, l2.x, a[c]);
                    a[c + 1] = fma(-lij, l2.y, a[c + 1]);
                }
            }
        }
        if (i == 0) { sm.bad = bad; sm.minp = minp; }
    }
    __syncthreads();
}

// ---- Q_blk = src D L^-T: forward substitution, one WARP per row, element c of the row in lane c ----------------------
// q_c = z_c / L[c][c] travels by shuffle, lane b > c takes z_b -= q_c L[b][c] (row b of L; zero above the diagonal, so no branch).
// A rolled loop of 32 steps: the one-thread-per-row version kept the row in registers, fully unrolled -- 2.8 k cycles per
// call and 13 k on the first one (17 KB of code fetched cold on every launch).
__device__ __forceinline__ void apply_phase(const PersistArgs& a, const double (*src)[PS], Smem& sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < a.RB; r += PT / 32) {
        double z = src[r][lane] * sm.dscale[lane];
        const double invd = sm.L[lane][lane];                                    // the diagonal holds 1 / L[c][c]
#pragma unroll 8
        for (int c = 0; c < PB; ++c) {
            const double lc = sm.L[lane][c];                                     // zero above the diagonal: lanes < c keep z
            const double q = __shfl_sync(0xffffffffu, z * invd, c);
            z = lane == c ? q : fma(-q, lc, z);
        }
        __syncwarp();
        sm.Q[r][lane] = z;
    }
    __syncthreads();
}

__device__ __forceinline__ void store_rows(const PersistArgs& a, double* __restrict__ Qdst, const double (*src)[PS], int r0) {
    for (int e = threadIdx.x; e < a.RB * PB; e += PT) {
        const int r = e / PB, c = e - r * PB;
        if (r0 + r < a.R) __stcg(Qdst + (int64_t)(r0 + r) * PB + c, src[r][c]);
    }
}

// ---- H (sm.A, symmetric B x B) = V diag(theta) V': parallel cyclic Jacobi, round-robin pairing -------------------------
// Result: sm.V = V with its columns sorted by descending theta, sm.th = theta.  One CTA; sm.A and sm.L are the two copies of
// the matrix, one CTA barrier per round.
//
// The phase is bound by shared-memory bandwidth, not by the rotation's latency chain (scripts/jacobi_probe.cu: with A and V
// as 32 x 33 arrays indexed by the pair of the round, a round moved 39 KB through a 128 B/clk pipe with 2-4-way bank conflicts
// on the pair-indexed accesses: 950-1,190 cycles, 150-184 k cycles for the five sweeps of C3).  So the matrix lives in "chair"
// order, as in a Brent-Luk systolic array: the 32 indices sit on chairs T[0..15] (top row) and B[0..15] (bottom row), pair i is
// (T[i], B[i]) in EVERY round, and after a round the values move one chair round the circle (T[0] stays):
//     T[i] -> T[i+1] (1 <= i <= 14),  T[15] -> B[15],  B[i] -> B[i-1] (i >= 1),  B[0] -> T[1].
// Thread (ia, ib) always owns rows (T[ia], B[ia]) x columns (T[ib], B[ib]):
//   * A is four 16 x 16 planes (row side, column side); the thread reads plane[.][.][ia][ib] -- lanes of a half-warp are
//     consecutive in ib, no bank conflict -- and writes each rotated entry to the chair it moves to (ib +- 1: again consecutive);
//   * the rotation of pair ib comes from the diagonal entries plane[.][.][ib][ib] (pitch 16: ib * 17, conflict-free), that of
//     pair ia by shuffle from lane ia (lane l holds pair l % 16);
//   * V (rows 2 ia, 2 ia + 1 x the two chairs of pair ib) stays in REGISTERS: its columns move to the neighbouring lane of the
//     same half-warp by shuffle, it never touches shared memory until the end.
// 1 / sqrt via rsqrt_fast, the rotation computed unconditionally and selected (no branch in the round).
#ifndef RUVIII_JSTAMP
#define RUVIII_JSTAMP(k)          // scripts/jacobi_probe.cu defines these to time the sections of a round
#define RUVIII_JSTAMP_DECL
#define RUVIII_JSTAMP_FLUSH
#endif

__device__ __forceinline__ int jacobi_phase(Smem& sm, int jorder = 0) {
    constexpr int H = PB / 2;                                    // pairs (16)
    constexpr int PL = H * H;                                    // doubles per plane
    static_assert(4 * PL <= PB * PS, "the four planes must fit one B x (B + 1) array");
    const int tid = threadIdx.x;
    const int ia = tid / H, ib = tid % H;
    double* const buf0 = &sm.A[0][0];
    constexpr int dAL = (int)((offsetof(Smem, L) - offsetof(Smem, A)) / sizeof(double));
    RUVIII_JSTAMP_DECL
    // chair (side s, position i) starts with index idx(s, i); plane (sr, sc) at (2 sr + sc) * PL
    auto idx = [jorder](int s_, int i) { return jorder == 1 ? (s_ ? PB - 1 - i : i) : jorder == 2 ? (s_ ? H + i : i) : 2 * i + s_; };
    const int ra0 = idx(0, ia), ra1 = idx(1, ia), cb0 = idx(0, ib), cb1 = idx(1, ib);
    {
        const double h00 = sm.A[ra0][cb0], h01 = sm.A[ra0][cb1];
        const double h10 = sm.A[ra1][cb0], h11 = sm.A[ra1][cb1];
        double* b1 = buf0 + dAL;
        b1[0 * PL + ia * H + ib] = h00;
        b1[1 * PL + ia * H + ib] = h01;
        b1[2 * PL + ia * H + ib] = h10;
        b1[3 * PL + ia * H + ib] = h11;
    }
    // V = I in chair order: v[u][s] = V[row 2 ia + u][chair (s, ib)]
    double vt0 = (2 * ia == cb0) ? 1.0 : 0.0, vb0 = (2 * ia == cb1) ? 1.0 : 0.0;
    double vt1 = (2 * ia + 1 == cb0) ? 1.0 : 0.0, vb1 = (2 * ia + 1 == cb1) ? 1.0 : 0.0;
    // where this thread's four entries move: next chair of T[i] / B[i] as (side, index)
    const int ntS_a = ia == H - 1 ? 1 : 0, ntI_a = ia == 0 ? 0 : (ia == H - 1 ? H - 1 : ia + 1);
    const int nbS_a = ia == 0 ? 0 : 1,     nbI_a = ia == 0 ? 1 : ia - 1;
    const int ntS_b = ib == H - 1 ? 1 : 0, ntI_b = ib == 0 ? 0 : (ib == H - 1 ? H - 1 : ib + 1);
    const int nbS_b = ib == 0 ? 0 : 1,     nbI_b = ib == 0 ? 1 : ib - 1;
    const int d00 = (2 * ntS_a + ntS_b) * PL + ntI_a * H + ntI_b;      // (T[ia], T[ib])
    const int d01 = (2 * ntS_a + nbS_b) * PL + ntI_a * H + nbI_b;      // (T[ia], B[ib])
    const int d10 = (2 * nbS_a + ntS_b) * PL + nbI_a * H + ntI_b;      // (B[ia], T[ib])
    const int d11 = (2 * nbS_a + nbS_b) * PL + nbI_a * H + nbI_b;      // (B[ia], B[ib])
    const int own = ia * H + ib, diag = ib * (H + 1);
    __syncthreads();
    int cur = 1, sweeps_run = 0;
    for (int sweep = 0; sweep < 16; ++sweep) {
        ++sweeps_run;
        if (tid == 0) { sm.flag = 0; sm.flag_big = 0; }
        int any = 0, big = 0;
        for (int round = 0; round < PB - 1; ++round) {
            RUVIII_JSTAMP(0);
            const double* A = buf0 + cur * dAL;
            double* An = buf0 + (cur ^ 1) * dAL;
            const double app = A[diag], apq = A[PL + diag], aqq = A[3 * PL + diag];
            const double x00 = A[own], x01 = A[PL + own], x10 = A[2 * PL + own], x11 = A[3 * PL + own];
            // cos(2 theta) = |zeta| / rr, |theta| <= pi / 4, two rsqrt and no division
            const double zeta = aqq - app, beta = 2.0 * apq;
            const double n2 = fma(zeta, zeta, beta * beta);
            const double lim = fabs(app * aqq), q2 = apq * apq;
            const bool rot = q2 > 1e-33 * lim && n2 > 1e-280;
            RUVIII_JSTAMP(1);
            const double irr = rsqrt_fast(n2);
            const double c2 = fma(0.5 * fabs(zeta), irr, 0.5);
            const double ic = rsqrt_fast(c2);
            const double cb = rot ? c2 * ic : 1.0;
            const double sb = rot ? (zeta >= 0.0 ? 0.5 : -0.5) * beta * irr * ic : 0.0;
            any |= rot;
            big |= rot && q2 > 1e-20 * lim;
            RUVIII_JSTAMP(2);
            const double ca = __shfl_sync(0xffffffffu, cb, ia), sa = __shfl_sync(0xffffffffu, sb, ia);
            RUVIII_JSTAMP(3);
            const double r00 = ca * x00 - sa * x10, r01 = ca * x01 - sa * x11;      // rows:    J_a' X
            const double r10 = sa * x00 + ca * x10, r11 = sa * x01 + ca * x11;
            An[d00] = cb * r00 - sb * r01;                                           // columns: (.) J_b, stored one chair on
            An[d01] = sb * r00 + cb * r01;
            An[d10] = cb * r10 - sb * r11;
            An[d11] = sb * r10 + cb * r11;
            // V <- V J_b, then the columns move with their chairs: to lane + 1 goes the top value (from lane 0: the bottom
            // one, B[0] -> T[1]); to lane - 1 the bottom value; T[0] stays and T[15] -> B[15] are the lane's own values
            {
                const double t0 = cb * vt0 - sb * vb0, b0 = sb * vt0 + cb * vb0;
                const double t1 = cb * vt1 - sb * vb1, b1 = sb * vt1 + cb * vb1;
                const double fl0 = __shfl_up_sync(0xffffffffu, ib == 0 ? b0 : t0, 1, H);
                const double fl1 = __shfl_up_sync(0xffffffffu, ib == 0 ? b1 : t1, 1, H);
                const double fr0 = __shfl_down_sync(0xffffffffu, b0, 1, H);
                const double fr1 = __shfl_down_sync(0xffffffffu, b1, 1, H);
                vt0 = ib == 0 ? t0 : fl0;  vb0 = ib == H - 1 ? t0 : fr0;
                vt1 = ib == 0 ? t1 : fl1;  vb1 = ib == H - 1 ? t1 : fr1;
            }
            cur ^= 1;
            RUVIII_JSTAMP(4);
            __syncthreads();
            RUVIII_JSTAMP(5);
        }
        if (ia == 0) {                                    // the 16 threads ia == 0 see every pair (ib = pair)
            if (any) sm.flag = 1;
            if (big) sm.flag_big = 1;
        }
        __syncthreads();
        // quadratic convergence: a sweep whose largest relative off-diagonal entry was below 1e-10 leaves them below rounding
        const bool stop = !sm.flag || !sm.flag_big;
        __syncthreads();
        if (stop) break;
    }
    RUVIII_JSTAMP_FLUSH
    // sort descending (rank sort over the chairs): sm.th, sm.V = V with its columns permuted
    const double* A = buf0 + cur * dAL;
    int* rank_of_chair = &sm.pr[0][0];
    if (tid < PB) {
        const double t = A[(tid < H ? 0 : 3 * PL) + (tid % H) * (H + 1)];
        int rank = 0;
        for (int c = 0; c < PB; ++c) {
            const double u = A[(c < H ? 0 : 3 * PL) + (c % H) * (H + 1)];
            rank += (u > t) || (u == t && c < tid);
        }
        rank_of_chair[tid] = rank;
        sm.th[rank] = t;
    }
    __syncthreads();
    {
        const int rt = rank_of_chair[ib], rb = rank_of_chair[H + ib];
        sm.V[2 * ia][rt] = vt0;      sm.V[2 * ia][rb] = vb0;
        sm.V[2 * ia + 1][rt] = vt1;  sm.V[2 * ia + 1][rb] = vb1;
    }
    __syncthreads();
    return sweeps_run;
}

__device__ __forceinline__ void trace_mark(const PersistArgs& a, int& n, int phase) {
    if (a.trace && blockIdx.x == 0 && threadIdx.x == 0 && n < kMaxTrace) a.trace[n++] = ((long long)phase << 48) | (clock64() & 0xFFFFFFFFFFFFll);
}

// dst (B x B) = lhs' rhs or lhs rhs for B x B shared-memory operands (every thread 4 entries)
__device__ __forceinline__ void small_mm(double (*dst)[PS], const double (*lhs)[PS], const double (*rhs)[PS], bool lhs_t) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int e = threadIdx.x + u * PT, i = e / PB, j = e - i * PB;
        double s = 0.0;
#pragma unroll 8
        for (int q = 0; q < PB; ++q) s = fma(lhs_t ? lhs[q][i] : lhs[i][q], rhs[q][j], s);
        dst[i][j] = s;
    }
}

enum Phase { PH_START = 1, PH_PRODUCT, PH_PARTIALS, PH_BAR, PH_REDUCE, PH_CHOL, PH_APPLY, PH_JACOBI, PH_ROTATE, PH_CHECK, PH_CHEB, PH_END };

template <int RT, bool GS>
__global__ void __launch_bounds__(PT, 1) ek_persistent_kernel(PersistArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    double* gs = reinterpret_cast<double*>(smem_raw + ((sizeof(Smem) + 15) & ~size_t(15)));   // GS: RB x ldg doubles behind Smem
    const int tid = threadIdx.x, cta = blockIdx.x, R = a.R, RB = a.RB, P = a.P;
    const int r0 = cta * RB;
    unsigned target = 0;
    int ntrace = 0;
    int cur = 0;
    double* Qbuf[2] = {a.Q0, a.Q1};
    trace_mark(a, ntrace, PH_START);
    if (cta == 0 && tid == 0) a.state[5] = 0;
    auto gbar = [&]() {
        grid_barrier(a.bar, target, P);
        trace_mark(a, ntrace, PH_BAR);
    };

    // One round of column-scaled Cholesky QR of the rows in `src` (partials -> reduce -> Cholesky -> substitution);
    // with `have_S` the reduced S is already in sm.S.  Leaves Q_blk in sm.Q; returns false on a Cholesky breakdown.
    double first_minp = 0.0;                                  // smallest pivot of the first round of the last ortho()
    auto ortho = [&](const double (*src)[PS], bool have_S, bool accurate) -> bool {
        for (int round = 0; round < 2; ++round) {
            if (!(round == 0 && have_S)) {
                partials_phase(a, round == 0 ? src : sm.Q, false, sm);
                trace_mark(a, ntrace, PH_PARTIALS);
                gbar();
                reduce_phase(a, 1024, sm);
                trace_mark(a, ntrace, PH_REDUCE);
                gbar();
                for (int e = tid; e < PB * PB; e += PT) sm.S[e / PB][e % PB] = __ldcg(a.Pred + e);
                __syncthreads();
            }
            cholesky_phase(sm);
            trace_mark(a, ntrace, PH_CHOL);
            if (sm.bad) return false;
            apply_phase(a, round == 0 ? src : sm.Q, sm);
            trace_mark(a, ntrace, PH_APPLY);
            // kappa^2 of the scaled block ~ 1 / min pivot: below 1e-4 the orthogonality error of one round exceeds ~1e-12,
            // so the block is orthogonalised once more (Cholesky QR 2); every CTA sees the same pivots and takes the same branch.
            // A block that only feeds the next product (not a Rayleigh-Ritz step) may be 1e-6 away from orthonormal.
            if (round == 0) first_minp = sm.minp;
            if (sm.minp >= (accurate ? 1e-4 : 1e-10)) break;
            if (round == 0 && cta == 0 && tid == 0) a.state[5] += 1;
        }
        return true;
    };

    // ---- start block: fixed hash of (row, column), uniform in (-1, 1).  It is not orthogonalised: the first product
    //      G Q_0 does not care, and its Cholesky QR (two rounds: the columns of G Q_0 are nearly parallel) does the job ----
    for (int e = tid; e < RB * PB; e += PT) {
        const int r = e / PB, c = e - r * PB;
        double v = 0.0;
        if (r0 + r < R) {
            uint64_t x = (uint64_t)((int64_t)(r0 + r) * PB + c) * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
            x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
            v = (double)(int64_t)(x >> 11) * (1.0 / 4503599627370496.0) - 1.0;
        }
        sm.Q[r][c] = v;
    }
    if (GS) {
        // this CTA's rows of G stay in shared memory for the whole solve (RB = 8: 8 R doubles)
        const int ldg = a.ldg;
        for (int e = tid; e < RB * R; e += PT) {
            const int r = e / R, c = e - r * R;
            gs[r * ldg + c] = (r0 + r < R) ? __ldg(a.G + (int64_t)(r0 + r) * R + c) : 0.0;
        }
    }
    __syncthreads();
    bool ok = true;
    int converged = 0, it = 0, n_rr = 0, n_cheb = 0, n_skip = 0;
    bool skipped_last = true;                                 // the first pass always orthogonalises
    double last_minp = 0.0;
    store_rows(a, Qbuf[cur], sm.Q, r0);
    gbar();
    int next_rr = a.first;
    while (ok && it < a.max_iter) {
        ++it;
        product_phase<RT, GS>(a, Qbuf[cur], r0, sm, gs);
        trace_mark(a, ntrace, PH_PRODUCT);
        const bool rr = it >= next_rr;
        if (!rr) {
            // G (G Q) spans what G orth(G Q) spans: when the last Cholesky QR found the (column-scaled) block well conditioned
            // -- the block is close to invariant, G Q ~ Q Lambda -- the next product starts from the unnormalised rows and
            // the orthogonalisation (partials, two grid barriers, Cholesky, substitution: 23 k cycles at C3, more than the
            // product itself) is done once for two products.  Never twice in a row, never right before a Rayleigh-Ritz step.
            const bool skip = a.skip_minp > 0.0 && it >= 2 && !skipped_last && it + 1 < next_rr && last_minp >= a.skip_minp;
            if (skip) {
                store_rows(a, Qbuf[cur ^ 1], sm.Z, r0);
                gbar();
                cur ^= 1;
                skipped_last = true;
                ++n_skip;
                continue;
            }
            ok = ortho(sm.Z, false, it + 1 >= next_rr);
            last_minp = skipped_last ? sqrt(first_minp) : first_minp;      // per product
            skipped_last = false;
            if (!ok) break;
            store_rows(a, Qbuf[cur ^ 1], sm.Q, r0);
            gbar();
            cur ^= 1;
            continue;
        }
        // ---- Rayleigh-Ritz pass ----
        ++n_rr;
        partials_phase(a, sm.Z, true, sm);
        gbar();
        reduce_phase(a, 2048, sm);
        gbar();
        trace_mark(a, ntrace, PH_REDUCE);
        if (cta == 0) {
            for (int e = tid; e < PB * PB; e += PT) sm.L[e / PB][e % PB] = __ldcg(a.Pred + 1024 + e);
            __syncthreads();
            for (int e = tid; e < PB * PB; e += PT) {
                const int i = e / PB, j = e % PB;
                sm.A[i][j] = 0.5 * (sm.L[i][j] + sm.L[j][i]);
            }
            __syncthreads();
            const int sweeps = jacobi_phase(sm, a.jorder);
            for (int e = tid; e < PB * PB; e += PT) __stcg(a.Vm + e, sm.V[e / PB][e % PB]);
            if (tid < PB) __stcg(a.theta + tid, sm.th[tid]);
            if (tid == 0) a.state[3] = sweeps;
            trace_mark(a, ntrace, PH_JACOBI);
        }
        gbar();
        for (int e = tid; e < PB * PB; e += PT) {
            sm.V[e / PB][e % PB] = __ldcg(a.Vm + e);
            sm.S[e / PB][e % PB] = __ldcg(a.Pred + e);
        }
        if (tid < PB) sm.th[tid] = __ldcg(a.theta + tid);
        __syncthreads();
        // X_blk = Q_blk V, Z_blk <- Z_blk V (row-local), residual and sign partials
        {
            double xv[4], zv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = tid + u * PT, r = e / PB, c = e - r * PB;
                double x = 0.0, z = 0.0;
                if (r < RB) {
#pragma unroll 8
                    for (int q = 0; q < PB; ++q) {
                        x = fma(sm.Q[r][q], sm.V[q][c], x);
                        z = fma(sm.Z[r][q], sm.V[q][c], z);
                    }
                }
                xv[u] = x; zv[u] = z;
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = tid + u * PT, r = e / PB, c = e - r * PB;
                if (r < RB) { sm.X[r][c] = xv[u]; sm.Z[r][c] = zv[u]; }
            }
            __syncthreads();
            if (tid < PB) {
                double s = 0.0, best = 0.0, best_abs = -1.0;
                const double th = sm.th[tid];
                for (int r = 0; r < RB; ++r) {
                    if (r0 + r >= R) break;
                    const double x = sm.X[r][tid], d = sm.Z[r][tid] - th * x;
                    s = fma(d, d, s);
                    if (fabs(x) > best_abs) { best_abs = fabs(x); best = x; }
                }
                __stcg(a.Pr + (int64_t)cta * 64 + tid, s);
                __stcg(a.Pr + (int64_t)cta * 64 + 32 + tid, best);
            }
        }
        trace_mark(a, ntrace, PH_ROTATE);
        gbar();
        // every CTA folds the P partials itself, in the same order: the same decision everywhere
        {
            const int j = tid % PB, grp = tid / PB;          // 8 groups of CTAs
            double s = 0.0, best = 0.0, best_abs = -1.0;
#pragma unroll 4
            for (int p = grp; p < P; p += PT / PB) {
                s += __ldcg(a.Pr + (int64_t)p * 64 + j);
                const double x = __ldcg(a.Pr + (int64_t)p * 64 + 32 + j);
                const double ax = fabs(x);
                if (ax > best_abs) { best_abs = ax; best = x; }
            }
            sm.part[grp * PB + j] = s;
            sm.part[1024 + grp * PB + j] = best;
            __syncthreads();
            if (tid < PB) {
                double t = 0.0, b = 0.0, b_abs = -1.0;
                for (int q = 0; q < PT / PB; ++q) {
                    t += sm.part[q * PB + tid];
                    const double x = sm.part[1024 + q * PB + tid], ax = fabs(x);
                    if (ax > b_abs) { b_abs = ax; b = x; }
                }
                sm.red[tid] = sqrt(t);
                sm.sgn[tid] = b < 0.0 ? -1.0 : 1.0;          // sign that makes the largest-|component| positive
            }
            __syncthreads();
            if (tid == 0) {
                bool conv = true;
                const double lim = a.tol * fabs(sm.th[0]);
                for (int q = 0; q < a.kk; ++q) conv = conv && (sm.red[q] <= lim);
                sm.flag = conv;
            }
            __syncthreads();
            converged = sm.flag;
        }
        trace_mark(a, ntrace, PH_CHECK);
        if (converged) break;
        // ---- not yet: filtered continuation ----
        // Damp [0, b], b = smallest Ritz value of the block (<= lambda_32 by interlacing; whatever the block has not captured
        // lies at or below it and is amplified less than every wanted pair).  With t(lambda) = 2 lambda / b - 1 a degree-d
        // Chebyshev filter gains T_d(t(theta_k)) ~ cosh(d acosh t_k) on wanted pair k.  Rounds of degree <= dcap, each
        // followed by Cholesky QR only (no Jacobi), until the predicted gain on the worst wanted pair reaches the tolerance
        // (x 10 margin); then one product and the next Rayleigh-Ritz.  dcap bounds what one Cholesky QR (+ its second round)
        // has to take: the contamination of an unconverged column by the leading direction grows like
        // (b / theta_0) T_d(t(theta_0)).  Spectra with a huge spread (C3: theta_0 / b ~ 10^3) get dcap = 1, i.e. shifted
        // power passes; slowly decaying ones get the high degrees they need.
        const double bnd = fmax(sm.th[PB - 1], 1e-300), th0 = fabs(sm.th[0]);
        const double tk = 2.0 * sm.th[a.kk - 1] / bnd - 1.0, t0 = 2.0 * th0 / bnd - 1.0;
        double worst = 0.0;
        for (int q = 0; q < a.kk; ++q) worst = fmax(worst, sm.red[q]);
        if (!a.cheb || !(tk > 1.0 + 1e-6)) {
            // plain pass: S of the rotated block follows from the reduced S: S <- V' S V (no second round of partials)
            small_mm(sm.A, sm.S, sm.V, false);
            __syncthreads();
            small_mm(sm.L, sm.V, sm.A, true);
            __syncthreads();
            for (int e = tid; e < PB * PB; e += PT) sm.S[e / PB][e % PB] = sm.L[e / PB][e % PB];
            __syncthreads();
            ok = ortho(sm.Z, true, true);
            last_minp = first_minp; skipped_last = false;
            next_rr = it + a.period;
        } else {
            const double need = log(fmax(10.0 * worst / (a.tol * th0), 2.0));
            const double ak = acosh(tk), a0 = acosh(fmax(t0, 1.0 + 1e-6));
            const int dcap = max(1, (int)floor((log(1e7 * (t0 + 1.0) * 0.5) + log(2.0)) / a0));
            const double ce = 0.5 * bnd;                     // half-width and centre of [0, b]
            const double sig1 = ce / (th0 - ce);             // scaled so that p(theta_0) stays O(1)
            double gain = 0.0;
            for (int e = tid; e < RB * PB; e += PT) { const int r = e / PB, c = e - r * PB; sm.Q[r][c] = sm.X[r][c]; }
            __syncthreads();
            while (gain < need && it < a.max_iter) {
                const int d = (int)fmin(fmin((double)dcap, fmax(1.0, ceil((need - gain + log(2.0)) / ak))),
                                        fmin(40.0, (double)(a.max_iter - it)));
                store_rows(a, Qbuf[cur ^ 1], sm.Q, r0);      // start block of this round: Ritz vectors / the last Q
                gbar();
                cur ^= 1;
                // Y_1 = (sigma_1 / e) (G - c I) Y_0,  Y_{i+1} = (2 sigma_{i+1} / e) (G - c I) Y_i - sigma_i sigma_{i+1} Y_{i-1}:
                // the rows of Y_i live in sm.Q, those of Y_{i-1} in sm.X; one barrier per product
                double sig = sig1;
                for (int i = 1; i <= d; ++i) {
                    product_phase<RT, GS>(a, Qbuf[cur], r0, sm, gs);
                    const double sig_next = 1.0 / (2.0 / sig1 - sig);
                    for (int e = tid; e < RB * PB; e += PT) {
                        const int r = e / PB, c = e - r * PB;
                        const double y = sm.Q[r][c], gy = sm.Z[r][c] - ce * y;
                        const double yn = i == 1 ? (sig1 / ce) * gy : (2.0 * sig_next / ce) * gy - sig * sig_next * sm.X[r][c];
                        sm.X[r][c] = y;
                        sm.Q[r][c] = yn;
                    }
                    if (i > 1) sig = sig_next;
                    __syncthreads();
                    ++it; ++n_cheb;
                    if (i < d) {
                        store_rows(a, Qbuf[cur ^ 1], sm.Q, r0);
                        gbar();
                        cur ^= 1;
                    }
                    trace_mark(a, ntrace, PH_CHEB);
                }
                gain += log(cosh((double)d * ak));
                for (int e = tid; e < RB * PB; e += PT) { const int r = e / PB, c = e - r * PB; sm.Z[r][c] = sm.Q[r][c]; }
                __syncthreads();
                ok = ortho(sm.Z, false, true);
                if (!ok) break;
            }
            skipped_last = true;                             // no skipping on what the filter rounds left
            next_rr = it + 1;                                // Rayleigh-Ritz right after the next product
        }
        if (!ok) break;
        store_rows(a, Qbuf[cur ^ 1], sm.Q, r0);
        gbar();
        cur ^= 1;
    }
    // ---- results ----
    if (converged) {
        for (int e = tid; e < RB * a.kp; e += PT) {
            const int r = e / a.kp, j = e - r * a.kp;
            if (r0 + r < R) a.U[(int64_t)(r0 + r) * a.kp + j] = j < a.kk ? sm.sgn[j] * sm.X[r][j] : 0.0;
        }
        if (cta == 0 && tid < a.n_keep && tid < PB) a.evals[tid] = sm.th[tid];
    }
    if (cta == 0 && tid == 0) {
        a.state[0] = converged;
        a.state[1] = it;
        a.state[2] = ok ? 0 : (sm.bad ? sm.bad : -1);
        a.state[4] = n_rr;
        a.state[6] = n_cheb;
        a.state[7] = n_skip;
        if (a.flag_out) *a.flag_out = converged ? 0.0 : 1.0;
    }
    trace_mark(a, ntrace, PH_END);
    if (a.trace && cta == 0 && tid == 0) a.trace[kMaxTrace] = ntrace;
    // the last CTA out re-arms the counters for the next launch
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned old = atomicAdd(a.bar + 1, 1u);
        if (old == (unsigned)P - 1) { a.bar[0] = 0; a.bar[1] = 0; __threadfence(); }
    }
}

template <int RT, bool GS>
void launch_persist(const PersistArgs& a, size_t smem, cudaStream_t s) {
    // per device (a process may drive several): cheap enough to repeat on every launch
    RUVIII_CUDA_CHECK(cudaFuncSetAttribute(ek_persistent_kernel<RT, GS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    static const bool coop = [] { const char* v = std::getenv("RUVIII_EIG_COOP"); return !(v && *v == '0'); }();
    PersistArgs args = a;
    if (coop) {
        void* params[] = {&args};
        RUVIII_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)ek_persistent_kernel<RT, GS>, dim3(a.P), dim3(PT), params, smem, s));
    } else {
        ek_persistent_kernel<RT, GS><<<a.P, PT, smem, s>>>(args);
        RUVIII_LAUNCH_CHECK();
    }
}

}  // namespace

// Rows of G per CTA for the persistent solver, or 0 when the order does not fit one co-resident grid.
int topk_persist_rows(int R, int n_sms) {
    const char* pv = std::getenv("RUVIII_EIG_PERSIST");
    const int off = pv && *pv == '0' ? 1 : 0;
    if (off || R < 4 * PB) return 0;
    const char* fv = std::getenv("RUVIII_EIG_RB");          // experiments: force 8 / 16 / 32 rows per CTA
    const int forced = fv && *fv ? std::atoi(fv) : 0;
    for (int rb : {8, 16, 32})
        if ((R + rb - 1) / rb <= n_sms && (!forced || rb >= forced)) return rb;
    return 0;
}

size_t topk_persist_doubles(int n_sms) { return (size_t)n_sms * (2048 + 64) + 2048 + PB * PB + PB + (kMaxTrace + 2); }

// Enqueues the solve and returns at once.  h_state (pinned, >= 8 ints) receives the device state behind the kernel on the
// same stream; the caller reads it after its own synchronisation (TopkResult::pending) or asks for `sync`.
TopkResult topk_eigen_persistent(const double* G, int R, int kk, int kp, int n_keep, TopkWork& w, double* persist_ws, double* U,
                                 double* evals, double* flag_out, double tol, int max_iter, int n_sms, int* h_state, bool sync,
                                 cudaStream_t s) {
    TopkResult res;
    const int rb = topk_persist_rows(R, n_sms);
    if (!rb || kk > kTopkMaxK || n_keep > PB) throw Error(6, "topk_eigen_persistent: outside the supported range");
    PersistArgs a{};
    a.G = G; a.R = R; a.RB = rb; a.P = (R + rb - 1) / rb;
    a.Q0 = w.Q; a.Q1 = w.X;
    double* p = persist_ws;
    a.Pp = p; p += (size_t)n_sms * 2048;
    a.Pr = p; p += (size_t)n_sms * 64;
    a.Pred = p; p += 2048;
    a.Vm = p; p += PB * PB;
    a.theta = p; p += PB;
    static const bool trace_on = [] { const char* v = std::getenv("RUVIII_EIG_TRACE"); return v && *v == '1'; }();
    a.trace = trace_on ? reinterpret_cast<long long*>(p) : nullptr;
    a.U = U; a.evals = evals; a.flag_out = flag_out;
    a.kk = kk; a.kp = kp; a.n_keep = n_keep; a.tol = tol; a.max_iter = max_iter;
    static const int first_env = [] { const char* v = std::getenv("RUVIII_EIG_FIRST"); return v && *v ? std::atoi(v) : 0; }();
    static const int cheb_env = [] { const char* v = std::getenv("RUVIII_EIG_CHEB"); return v && *v ? std::atoi(v) : 1; }();
    a.first = std::max(2, first_env > 0 ? first_env : 6);   // the start block is not orthonormal: no Rayleigh-Ritz on pass 1
    a.period = 3;
    a.cheb = cheb_env;
    static const int jorder_env = [] { const char* v = std::getenv("RUVIII_EIG_JORDER"); return v && *v ? std::atoi(v) : 0; }();
    a.jorder = jorder_env;
    static const double skip_env = [] { const char* v = std::getenv("RUVIII_EIG_SKIP"); return v && *v ? std::atof(v) : 1e-3; }();
    a.skip_minp = skip_env;
    a.bar = reinterpret_cast<unsigned*>(w.state + 8);
    a.state = w.state;
    // RB = 8: the CTA's 8 rows of G (8 R doubles, pitch = 4 mod 16 doubles so that the 8 x 4 fragment loads of a warp spread
    // over the banks) live in shared memory behind the Smem block
    const size_t base_smem = (sizeof(Smem) + 15) & ~size_t(15);
    a.ldg = (int)round_up(R, 16) + 4;
    const size_t gs_smem = base_smem + (size_t)8 * a.ldg * sizeof(double);
    static const bool gs_off = [] { const char* v = std::getenv("RUVIII_EIG_GSMEM"); return v && *v == '0'; }();
    switch (rb) {
        case 8:
            if (!gs_off && gs_smem <= 200 * 1024) launch_persist<1, true>(a, gs_smem, s);
            else launch_persist<1, false>(a, base_smem, s);
            break;
        case 16: launch_persist<2, false>(a, base_smem, s); break;
        default: launch_persist<4, false>(a, base_smem, s); break;
    }
    res.launches = 1;
    RUVIII_CUDA_CHECK(cudaMemcpyAsync(h_state, w.state, 8 * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (!sync && !trace_on) {
        res.pending = true;
        res.converged = true;
        return res;
    }
    for (int spin = 0; spin < (1 << 22) && cudaStreamQuery(s) == cudaErrorNotReady; ++spin) {}
    RUVIII_CUDA_CHECK(cudaStreamSynchronize(s));
    res.converged = h_state[0] != 0;
    res.iterations = h_state[1];
    if (trace_on) {
        std::vector<long long> t(kMaxTrace + 2);
        RUVIII_CUDA_CHECK(cudaMemcpy(t.data(), a.trace, t.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        const int n = (int)std::min<long long>(t[kMaxTrace], kMaxTrace);
        static const char* names[] = {"?", "start", "product", "partials", "bar", "reduce", "chol", "apply", "jacobi", "rotate", "check",
                                      "cheb", "end"};
        fprintf(stderr, "[eig persist] R=%d RB=%d P=%d converged=%d products=%d rr=%d sweeps=%d chol2=%d cheb=%d skipped=%d; cycles:", R, rb, a.P,
                h_state[0], h_state[1], h_state[4], h_state[3], h_state[5], h_state[6], h_state[7]);
        for (int i = 1; i < n; ++i)
            fprintf(stderr, " %s=%lld", names[std::min<long long>(t[i] >> 48, 12)], (t[i] & 0xFFFFFFFFFFFFll) - (t[i - 1] & 0xFFFFFFFFFFFFll));
        fprintf(stderr, " total=%lld\n", n > 1 ? (t[n - 1] & 0xFFFFFFFFFFFFll) - (t[0] & 0xFFFFFFFFFFFFll) : 0);
    }
    return res;
}

}  // namespace ruviii
