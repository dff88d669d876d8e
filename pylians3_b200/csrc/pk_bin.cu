// K5/K6: MAS-window deconvolution + |delta_k|^2 binning of the r2c half-spectrum, one read pass.
//
// Replaces the serial loops library/Pk_library/Pk_library.pyx:311-378 (Pk) and :623-732 (XPk).
//
// Why not "one thread per mode + atomics": each mode feeds ~7 float64 accumulators (3D shell:
// k, P0, P2, P4, phase; 2D (k_par,k_per) bin; 1D k_par bin) plus three counters.  At the
// HBM roofline a SM must retire ~3 modes/clock, while red.global issues < 1 lane/clock/SM
// and shared-memory float64 atomics are CAS loops -- per-mode atomics cap the kernel at a
// few percent of the roofline.  So the kernel removes the atomics geometrically instead:
//
//  * WALK.  A thread owns a column of modes: fixed |k_o| (the "other" in-plane axis) and
//    fixed kz (lanes run along kz, the contiguous axis, so every warp load is a coalesced
//    256-byte row segment), and walks |k_w| = s0..s1 along the remaining axis.  Along the
//    walk k^2 = |k_o|^2 + kz^2 + s^2 grows monotonically, so the shell index, and (because the
//    walk axis is chosen different from the line of sight) the (k_par,k_per) bin move
//    monotonically and slowly: the thread keeps the current bin's sums in REGISTERS and only
//    issues red.global.add.f64 when the bin changes.  The 1D bin (k_par) is constant per
//    thread and is flushed once.
//  * FOLD.  The four modes (+-k_o, +-k_w, kz) share |k|, mu^2, k_par, k_per, hence every bin
//    and Legendre weight and the (even) MAS window: they are loaded together and summed
//    before anything else happens, so the geometry math and the flushes are paid once per
//    four modes.  The Hermitian-duplicate rule of the reference (:324-327) is applied per
//    mode, so counts stay exact.
//  * SHELL WINDOW.  Measured (profiles/r2_pkbin.md): with per-lane red.global flushes the kernel was bound by the
//    L2 atomic path -- 1% of its instructions, 58% of its stall samples -- because the 3D shell sums are few, hot
//    addresses.  Every warp therefore owns a private window of PK_WBINS shells in shared memory.  A lane whose
//    shell changes adds its register sums to the window with plain read-modify-writes; lanes of one warp that
//    leave the SAME shell in the same step are neighbours (k grows with kz along the lanes) and take turns.  The
//    window reaches global memory once per warp and walk segment, coalesced, through NREP replicas folded by a
//    second tiny kernel.  The 2D (k_par,k_per) array is large and cold: its flushes stay direct.
//
// Counts are uint64, everything else float64, exactly as wide as the reference's accumulators.
#include <math.h>

#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

namespace pyl {

constexpr int PK_BLOCK = 128;
constexpr int PK_NREP = 8;
constexpr int PK_WBINS = 96;          // shells a warp's private window covers: 31 (its kz span) + the walk segment
constexpr int PK_MAX_SEG = 60;        // hence at most this many walk steps per segment

template <int F>
struct PkArgs {
    const float2 *dk[F];     // (N, nky, nz) complex64 each
    const double *win[F];    // MAS window per |k| index: (x/sin x)^p, x = pi*k/N, [m+1]
    int N, m, nz, even;
    int ky_lo, nky;          // stored ky rows: nky rows, the first ones are global ky = ky_lo, ky_lo+1, ...
    int ny_lo, ky_up0;       // mirrored slab: rows [0,ny_lo) hold ky_lo.., rows [ny_lo,nky) hold global ky_up0..
                             // (the mirrors N-ky of the lower rows, ascending); otherwise ny_lo = nky
    int walk_y;              // 1: walk along ky, other axis = kx;  0: walk along kx, other = ky
    int fold_other;          // 1: thread folds +-k_o (needs both rows stored)
    int n_other;             // fold: number of |k_o| values; no fold: stored indices of the other axis
    int o_lo;                // fold: first |k_o| (0 unless the other axis is a mirrored ky window)
    int w_lo, w_hi;          // walk range |k_w| in [w_lo, w_hi)
    int axis;                // line of sight
    int ky_major;            // fields are (nky, N, nz) -- stored ky row outermost -- instead of (N, nky, nz)
    int cross_imag;          // cross term: 0 = re_i re_j + im_i im_j (XPk), 1 = im_i re_j - re_i im_j (XPk_imag)
    int kmax_par1;           // kmax_par + 1
    int seg_len, nseg;       // walk range [0, m] cut into nseg segments of seg_len steps
    int nzp;                 // nz rounded up to a multiple of 32: a warp holds 32 consecutive kz of ONE row
    long long T;             // threads per segment = n_other * nzp
    unsigned long long *out; // layout base (2D section is accumulated here directly)
    unsigned long long *rep; // NREP replicas of words [0, rep_words) (3D + 1D sections)
    long long rep_words;
    // word offsets (see pyl_pk_layout_t)
    long long o_k3D, o_Nm3D, o_Pk3D, o_PkX3D, o_phase, o_Nm1D, o_Pk1D, o_PkX1D, o_Nm2D, o_Pk2D, o_PkX2D;
};

__device__ __forceinline__ int isqrt_fix(int v) {
    int r = (int)__fsqrt_rn((float)v);
    if (r * r > v) r--;
    else if ((r + 1) * (r + 1) <= v) r++;
    return r;
}

__device__ __forceinline__ void red_f64(unsigned long long *base, long long word, double v) {
    atomicAdd(reinterpret_cast<double *>(base + word), v);   // result unused -> RED.E.ADD.F64
}
__device__ __forceinline__ void red_u64(unsigned long long *base, long long word, unsigned long long v) {
    atomicAdd(base + word, v);
}

// phase^2 of one mode, phase = atan2(re, |delta_k|) (Pk_library.pyx:358).  |delta_k| >= |re|,
// so |phase| = atan(t) with t = 1/sqrt(1 + (im/re)^2) in (0,1], evaluated in float32 (scale-free, no
// overflow of re^2+im^2).  atan on [0,1] is the 9-term odd polynomial of Abramowitz & Stegun 4.4.49
// (|error| <= 2e-8, an order of magnitude inside float32 rounding): a third of libm atanf's instructions, which
// were 13% of this kernel's.  The per-mode relative error ~1e-7 is far inside the 1e-4 bin tolerance.
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ double phase_sq(float re, float im) {
    if (re == 0.0f) return 0.0;
    const float q = im * rcp_approx(re);
    const float u = rcp_approx(fmaf(q, q, 1.0f));       // t^2, t = |re|/|delta_k|
    float p = 0.0028662257f;
    p = fmaf(p, u, -0.0161657367f);
    p = fmaf(p, u, 0.0429096138f);
    p = fmaf(p, u, -0.0752896400f);
    p = fmaf(p, u, 0.1065626393f);
    p = fmaf(p, u, -0.1420889944f);
    p = fmaf(p, u, 0.1999355085f);
    p = fmaf(p, u, -0.3333314528f);
    p = fmaf(p, u, 1.0f);
    return (double)(u * p * p);                           // (t * P(t^2))^2
}

// Register cap: tried and dropped.  Built with __launch_bounds__(128, 8) (64 registers, 32 warps/SM instead of 24)
// the kernel ran in the same 0.89 ms at 512^3 -- it is not bound by occupancy but by the LSU/L2 path of its
// red.global flushes and table loads (profiles/r1_pkbin_hotlines.md).
// GEOM = false: the mode counts and the sum of |k| per bin -- functions of the geometry alone -- are not
// accumulated (the caller restores them from its per-geometry cache): a third fewer red.global per flush and no
// float64 square root per step.
#ifndef PYL_PK_MINB
#define PYL_PK_MINB 1
#endif
template <int F, bool PHASE, bool GEOM>
__global__ void __launch_bounds__(PK_BLOCK, PYL_PK_MINB) pk_bin_walk_kernel(const PkArgs<F> A) {
    constexpr int X = F * (F - 1) / 2;
    const int seg = blockIdx.y;
    const long long t = (long long)blockIdx.x * PK_BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31;

    const int N = A.N, m = A.m, nz = A.nz;
    const bool even = A.even != 0;
    int oi = (int)(t / A.nzp);             // index along the other axis (|k_o| or stored index)
    int kz = (int)(t - (long long)oi * A.nzp);
    // lanes past the end of a row (or of the thread range) stay in the loop for the warp-wide steps below but
    // never load or accumulate anything
    const bool active = (t < A.T) && (kz < nz);
    if (!active) { oi = 0; kz = 0; }
    const bool zspecial = (kz == 0) || (kz == m && even);

    // this warp's private shell window (see the header comment)
    constexpr int NV = 3 * F + 3 * (F * (F - 1) / 2) + (PHASE ? 1 : 0);
    extern __shared__ double pk_window[];
    double *wwin = pk_window + (threadIdx.x >> 5) * (PK_WBINS * NV);
    for (int i = lane; i < PK_WBINS * NV; i += 32) wwin[i] = 0.0;
    __syncwarp();

    // ---- the (up to two) rows of the other axis handled by this thread ------------------
    int o_val[2], o_idx[2];
    bool o_ok[2];
    int a_abs;
    // stored row of a global ky index (identity on a whole grid)
    auto ystore = [&](int gy) -> int {
        return (gy - A.ky_lo < A.ny_lo && gy >= A.ky_lo) ? gy - A.ky_lo : A.ny_lo + (gy - A.ky_up0);
    };
    if (A.fold_other) {
        a_abs = A.o_lo + oi;
        o_val[0] = a_abs; o_idx[0] = a_abs; o_ok[0] = true;
        o_val[1] = -a_abs; o_idx[1] = N - a_abs;
        o_ok[1] = (a_abs != 0) && !(even && a_abs == m);
        if (!A.walk_y) {                     // the other axis is y: translate to stored rows
            o_idx[0] = ystore(o_idx[0]);
            if (o_ok[1]) o_idx[1] = ystore(o_idx[1]);
        }
    } else {
        const int gi = oi + (A.walk_y ? 0 : A.ky_lo);    // stored index -> global index
        const int v = (gi > m) ? gi - N : gi;
        a_abs = v < 0 ? -v : v;
        o_val[0] = v; o_idx[0] = A.walk_y ? gi : oi; o_ok[0] = true;
        o_val[1] = 0; o_idx[1] = 0; o_ok[1] = false;
    }

    // window factors that do not change along the walk
    double win_oz[F];
#pragma unroll
    for (int f = 0; f < F; f++) win_oz[f] = A.win[f][a_abs] * A.win[f][kz];

    // ---- bin bookkeeping ----------------------------------------------------------------
    const int s0 = A.w_lo + seg * A.seg_len;
    const int s1 = min(s0 + A.seg_len, A.w_hi);
    const int base2 = a_abs * a_abs + kz * kz;
    // k_par by line of sight; the walk coordinate is x (walk_y==0) or y (walk_y==1)
    const int walk_axis = A.walk_y ? 1 : 0;
    const int other_axis = A.walk_y ? 0 : 1;
    const bool par_is_walk = (A.axis == walk_axis);
    const int kpar_fixed = (A.axis == other_axis) ? a_abs : kz;     // used when !par_is_walk
    const int perp_base = par_is_walk ? base2 : (A.axis == other_axis ? kz * kz : a_abs * a_abs);

    int kidx = isqrt_fix(base2 + s0 * s0);
    // shell of lane 0 at the first step of the segment: the base of the warp's window (k grows with kz and s)
    const int bin0 = __shfl_sync(0xffffffffu, kidx, 0);
    int kper = isqrt_fix(perp_base + (par_is_walk ? 0 : s0 * s0));
    const int mm = m * m;

    int cur3 = -1, cur1 = -1, cur2 = -1;
    unsigned int cnt3 = 0, cnt2 = 0, cnt1 = 0;
    double ksum = 0.0, ph3 = 0.0;
    double P3[3][F], P2[F], P1[F];
    double PX3[3][X > 0 ? X : 1], PX2[X > 0 ? X : 1], PX1[X > 0 ? X : 1];
#pragma unroll
    for (int f = 0; f < F; f++) { P3[0][f] = P3[1][f] = P3[2][f] = 0.0; P2[f] = 0.0; P1[f] = 0.0; }
#pragma unroll
    for (int x = 0; x < X; x++) { PX3[0][x] = PX3[1][x] = PX3[2][x] = 0.0; PX2[x] = 0.0; PX1[x] = 0.0; }

    unsigned long long *rep = A.rep + (long long)(blockIdx.x % PK_NREP) * A.rep_words;

    // `leaving`: this lane's shell changed (or the walk ended) and it holds sums for the shell it leaves.  Warp-wide.
    auto flush3 = [&](bool leaving) {
        leaving = leaving && cnt3 != 0;
        const unsigned fl = __ballot_sync(0xffffffffu, leaving);
        if (fl == 0) return;
        if (GEOM && leaving) {                         // first call for a geometry only
            red_u64(rep, A.o_Nm3D + cur3, cnt3);
            red_f64(rep, A.o_k3D + cur3, ksum);
        }
        // lanes leaving the same shell are neighbours: they take turns, in lane order
        const int up = __shfl_up_sync(0xffffffffu, cur3, 1);
        const bool head = leaving && !(lane > 0 && ((fl >> (lane - 1)) & 1u) && up == cur3);
        const unsigned heads = __ballot_sync(0xffffffffu, head);
        const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
        const int turn = leaving ? lane - start : 0;
        const int turns = __reduce_max_sync(0xffffffffu, turn);
        const int wi = cur3 - bin0;
        const bool inwin = wi >= 0 && wi < PK_WBINS;
        double *cell = wwin + (inwin ? wi : 0) * NV;
        for (int r = 0; r <= turns; r++) {
            if (leaving && turn == r) {
                if (inwin) {
#pragma unroll
                    for (int l = 0; l < 3; l++) {
#pragma unroll
                        for (int f = 0; f < F; f++) cell[l * F + f] += P3[l][f];
#pragma unroll
                        for (int x = 0; x < X; x++) cell[3 * F + l * X + x] += PX3[l][x];
                    }
                    if (PHASE) cell[NV - 1] += ph3;
                } else {                               // outside the window (cannot happen for seg_len <= PK_MAX_SEG)
#pragma unroll
                    for (int l = 0; l < 3; l++) {
#pragma unroll
                        for (int f = 0; f < F; f++) red_f64(rep, A.o_Pk3D + ((long long)cur3 * 3 + l) * F + f, P3[l][f]);
#pragma unroll
                        for (int x = 0; x < X; x++) red_f64(rep, A.o_PkX3D + ((long long)cur3 * 3 + l) * X + x, PX3[l][x]);
                    }
                    if (PHASE) red_f64(rep, A.o_phase + cur3, ph3);
                }
            }
            __syncwarp();
        }
        if (leaving) {
            cnt3 = 0; ksum = 0.0; ph3 = 0.0;
#pragma unroll
            for (int f = 0; f < F; f++) P3[0][f] = P3[1][f] = P3[2][f] = 0.0;
#pragma unroll
            for (int x = 0; x < X; x++) PX3[0][x] = PX3[1][x] = PX3[2][x] = 0.0;
        }
    };
    auto flush2 = [&]() {
        if (cnt2 == 0) return;
        if (GEOM) red_u64(A.out, A.o_Nm2D + (long long)cur2, cnt2);
#pragma unroll
        for (int f = 0; f < F; f++) red_f64(A.out, A.o_Pk2D + (long long)cur2 * F + f, P2[f]);
#pragma unroll
        for (int x = 0; x < X; x++) red_f64(A.out, A.o_PkX2D + (long long)cur2 * X + x, PX2[x]);
        cnt2 = 0;
#pragma unroll
        for (int f = 0; f < F; f++) P2[f] = 0.0;
#pragma unroll
        for (int x = 0; x < X; x++) PX2[x] = 0.0;
    };
    auto flush1 = [&]() {
        if (cnt1 == 0) return;
        if (GEOM) red_u64(rep, A.o_Nm1D + cur1, cnt1);
#pragma unroll
        for (int f = 0; f < F; f++) red_f64(rep, A.o_Pk1D + (long long)cur1 * F + f, P1[f]);
#pragma unroll
        for (int x = 0; x < X; x++) red_f64(rep, A.o_PkX1D + (long long)cur1 * X + x, PX1[x]);
        cnt1 = 0;
#pragma unroll
        for (int f = 0; f < F; f++) P1[f] = 0.0;
#pragma unroll
        for (int x = 0; x < X; x++) PX1[x] = 0.0;
    };

    // ---- loads: the 2x2 sign combinations of (other, walk) at walk step s ----------------
    // slot c = 2*io + iw.  The element of slot c at step s sits at off[c] + s*dstep[c] (affine in s: the walk index is
    // s or N-s, and stored ky rows are contiguous in both halves of a mirrored slab), so the loop advances four
    // offsets instead of rebuilding four 64-bit addresses.  ok carries existence AND the Hermitian-duplicate rule;
    // it depends on s only at s = 0 and at the Nyquist step, which take the generic path.
    // o_index is a kx index (walk_y) or an already translated stored ky row (walk x)
    auto row_of = [&](int o_index, int w_index) -> long long {
        const int kxx = A.walk_y ? o_index : w_index;
        const int yrow = A.walk_y ? ystore(w_index) : o_index;
        return A.ky_major ? ((long long)yrow * N + kxx) * nz + kz : ((long long)kxx * A.nky + yrow) * nz + kz;
    };
    auto mode_ok = [&](int ov, int wv) -> bool {
        const int kx = A.walk_y ? ov : wv;
        const int ky = A.walk_y ? wv : ov;
        if (zspecial) {
            if (kx < 0) return false;
            if ((kx == 0 || (kx == m && even)) && ky < 0) return false;
        }
        return true;
    };
    auto ok_mask = [&](int s) -> unsigned {
        unsigned mask = 0;
        const bool wneg = (s != 0) && !(even && s == m);
#pragma unroll
        for (int io = 0; io < 2; io++)
#pragma unroll
            for (int iw = 0; iw < 2; iw++)
                if (active && o_ok[io] && (iw == 0 || wneg) && mode_ok(o_val[io], iw ? -s : s)) mask |= 1u << (2 * io + iw);
        return mask;
    };
    const int wstride = A.ky_major ? (A.walk_y ? N * nz : nz) : (A.walk_y ? nz : A.nky * nz);
    long long off[4];
    {
        const int sref = s0 > 0 ? s0 : 1;               // N - s is a stored index for s >= 1
#pragma unroll
        for (int io = 0; io < 2; io++) {
            off[2 * io] = row_of(o_idx[io], s0);
            off[2 * io + 1] = row_of(o_idx[io], N - sref) + (long long)(sref - s0) * wstride;
        }
    }
    const unsigned mask_mid = ok_mask(1 < m || !even ? 1 : 0);         // any interior step: s != 0, s != Nyquist
    auto step_mask = [&](int s) -> unsigned {
        return (s == 0 || (even && s == m)) ? ok_mask(s) : mask_mid;
    };
    auto fetch = [&](unsigned okmask, float2 (&v)[4][F]) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            if (okmask & (1u << c)) {
#pragma unroll
                for (int f = 0; f < F; f++) v[c][f] = __ldg(A.dk[f] + off[c]);
            }
            off[c] += (c & 1) ? -wstride : wstride;
        }
    };

    float2 cur_v[4][F];
    unsigned cur_ok = 0;
    double cur_w[F];                       // window factor of the walk coordinate, fetched one step ahead like the rows
    if (s0 < s1) {
        cur_ok = step_mask(s0);
        fetch(cur_ok, cur_v);
#pragma unroll
        for (int f = 0; f < F; f++) cur_w[f] = A.win[f][s0];
    }

    for (int s = s0; s < s1; s++) {
        // software prefetch of the next step while this one is reduced (the table load, issued at its point of
        // use, was 14% of the kernel's stall samples: profiles/r1_pkbin_hotlines.md)
        float2 nxt_v[4][F];
        unsigned nxt_ok = 0;
        double nxt_w[F];
#pragma unroll
        for (int f = 0; f < F; f++) nxt_w[f] = cur_w[f];
        if (s + 1 < s1) {
            nxt_ok = step_mask(s + 1);
            fetch(nxt_ok, nxt_v);
#pragma unroll
            for (int f = 0; f < F; f++) nxt_w[f] = A.win[f][s + 1];
        }

        // ---- geometry of this step (shared by the folded modes) -------------------------
        // one step raises k by less than one: the shell index (and k_per) moves by at most one
        const int ss = s * s;
        const int k2 = base2 + ss;
        if ((kidx + 1) * (kidx + 1) <= k2) kidx++;
        if (!par_is_walk) {
            const int p2 = perp_base + ss;
            if ((kper + 1) * (kper + 1) <= p2) kper++;
        }
        const int kpar = par_is_walk ? s : kpar_fixed;
        const int i2 = A.kmax_par1 * kper + kpar;
        const int i1 = (k2 <= mm) ? kpar : -1;

        flush3(kidx != cur3);
        cur3 = kidx;
        if (i2 != cur2) { flush2(); cur2 = i2; }
        if (i1 != cur1) { flush1(); cur1 = i1; }

        const int mult = __popc(cur_ok);
        if (mult) {
            // ---- window-deconvolved power of the folded modes ---------------------------
            double D[F], DX[X > 0 ? X : 1];
            double PH = 0.0;
#pragma unroll
            for (int f = 0; f < F; f++) D[f] = 0.0;
#pragma unroll
            for (int x = 0; x < X; x++) DX[x] = 0.0;
            float fac[F];
#pragma unroll
            for (int f = 0; f < F; f++)   // product in float64, rounded to float32 (:351)
                fac[f] = __double2float_rn(win_oz[f] * cur_w[f]);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                if (cur_ok & (1u << c)) {
                    double re[F], im[F];
#pragma unroll
                    for (int f = 0; f < F; f++) {
                        const float r32 = __fmul_rn(cur_v[c][f].x, fac[f]);   // complex64*float32 (:352)
                        const float i32 = __fmul_rn(cur_v[c][f].y, fac[f]);
                        re[f] = (double)r32;
                        im[f] = (double)i32;
                        D[f] += re[f] * re[f] + im[f] * im[f];
                        if (PHASE && f == 0) PH += phase_sq(r32, i32);
                    }
                    int x = 0;
#pragma unroll
                    for (int i = 0; i < F; i++)
#pragma unroll
                        for (int j = i + 1; j < F; j++) {
                            // Pk_library.pyx:716-717 (XPk) or :1000-1001 (XPk_imag)
                            DX[x] += A.cross_imag ? im[i] * re[j] - re[i] * im[j] : re[i] * re[j] + im[i] * im[j];
                            x++;
                        }
                }
            }

            // ---- Legendre weights (:347-348, :374-376) ----------------------------------
            // mu^2 = k_par^2 / k^2: reciprocal of the integer k^2 by one Newton step from the float32 seed --
            // relative error ~4e-15, a float64 division costs four times as much
            double mu2 = 0.0;
            if (k2 != 0) {
                const double dk2 = (double)k2;
                double r = (double)__frcp_rn((float)k2);
                r = r * (2.0 - dk2 * r);
                mu2 = (double)(kpar * kpar) * r;
            }
            const double val1 = (3.0 * mu2 - 1.0) * 0.5;
            const double val2 = (35.0 * mu2 * mu2 - 30.0 * mu2 + 3.0) * 0.125;

            cnt3 += mult;
            if (GEOM) ksum += (double)mult * sqrt((double)k2);
            cnt2 += mult;
            if (PHASE) ph3 += PH;
#pragma unroll
            for (int f = 0; f < F; f++) {
                P3[0][f] += D[f]; P3[1][f] += D[f] * val1; P3[2][f] += D[f] * val2;
                P2[f] += D[f];
            }
#pragma unroll
            for (int x = 0; x < X; x++) {
                PX3[0][x] += DX[x]; PX3[1][x] += DX[x] * val1; PX3[2][x] += DX[x] * val2;
                PX2[x] += DX[x];
            }
            if (i1 >= 0) {
                cnt1 += mult;
#pragma unroll
                for (int f = 0; f < F; f++) P1[f] += D[f];
#pragma unroll
                for (int x = 0; x < X; x++) PX1[x] += DX[x];
            }
        }

        cur_ok = nxt_ok;
#pragma unroll
        for (int f = 0; f < F; f++) cur_w[f] = nxt_w[f];
#pragma unroll
        for (int c = 0; c < 4; c++)
#pragma unroll
            for (int f = 0; f < F; f++) cur_v[c][f] = nxt_v[c][f];
    }
    flush3(true); flush2(); flush1();

    // ---- the warp's window -> global replicas, consecutive lanes on consecutive shells ---------------------------
    __syncwarp();
    for (int i = lane; i < PK_WBINS * NV; i += 32) {
        const int b = i % PK_WBINS, v = i / PK_WBINS;         // lanes run over shells for one value at a time
        const double val = wwin[b * NV + v];
        if (val != 0.0) {
            const long long bin = bin0 + b;
            long long word;
            if (v < 3 * F) word = A.o_Pk3D + (bin * 3 + v / F) * F + v % F;
            else if (v < 3 * F + 3 * X) word = A.o_PkX3D + (bin * 3 + (v - 3 * F) / (X > 0 ? X : 1)) * X + (v - 3 * F) % (X > 0 ? X : 1);
            else word = A.o_phase + bin;
            red_f64(rep, word, val);
        }
    }
}

// window table: win[f][i] = (x/sin x)^p, x = pi*i/N  (Pk_library.pyx:83-84; even in k)
__global__ void pk_window_kernel(double *tab, int fields, int m1, int N, const int4 p4) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m1) return;
    const int p[4] = {p4.x, p4.y, p4.z, p4.w};
    const double x = (M_PI / (double)N) * (double)i;
    for (int f = 0; f < fields; f++) {
        double v = 1.0;
        if (i != 0 && p[f] != 0) v = pow(x / sin(x), (double)p[f]);
        tab[(long long)f * m1 + i] = v;
    }
}

// out[w] = sum over replicas; words in [c0,c0+nc) and [c1,c1+nc1) are uint64 counts
__global__ void pk_fold_replicas_kernel(unsigned long long *out, const unsigned long long *rep,
                                        long long words, int nrep, long long c0, long long nc0,
                                        long long c1, long long nc1) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= words) return;
    const bool is_count = (w >= c0 && w < c0 + nc0) || (w >= c1 && w < c1 + nc1);
    if (is_count) {
        unsigned long long s = 0;
        for (int r = 0; r < nrep; r++) s += rep[(long long)r * words + w];
        out[w] = s;
    } else {
        double s = 0.0;
        for (int r = 0; r < nrep; r++) s += __longlong_as_double((long long)rep[(long long)r * words + w]);
        out[w] = (unsigned long long)__double_as_longlong(s);
    }
}

// uint64 counts -> float64 in place (exact below 2^53): lets ONE float64 sum all-reduce cover the whole accumulator
// block of a slab-distributed spectrum
__global__ void pk_counts_to_f64_kernel(unsigned long long *acc, long long c0, long long n0, long long c1, long long n1,
                                        long long c2, long long n2) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long w = -1;
    if (t < n0) w = c0 + t;
    else if (t < n0 + n1) w = c1 + (t - n0);
    else if (t < n0 + n1 + n2) w = c2 + (t - n0 - n1);
    if (w >= 0) acc[w] = (unsigned long long)__double_as_longlong((double)acc[w]);
}

// ---- finalisation on the device (Pk_library.pyx:384-418 / :735-791): units, averages, (2l+1) factors ----------
// In place on the accumulator buffer: every word becomes the float64 value the reference stores in the same
// slot (counts become float64 too; the DC bins keep their slots and are dropped by the caller).  One thread per
// bin; each expression is spelled with explicit-rounding intrinsics in the order of the reference's (and of
// Pk_library._finalize's) NumPy expressions, so the result is bit-identical to the host finalisation.
struct FinalizeArgs {
    unsigned long long *acc;
    double *kpar, *kper;     // optional: bin-centre coordinates of the 2D array (Pk_library.pyx:394-399)
    int kmax_par1;
    int F, X, n3, n1;
    long long n2;
    int counts_f64;
    double kF, kN2, fact, twopi2;   // kN2, fact, twopi2 come from libm pow() like the Python-level expressions
    long long o_k3D, o_Nm3D, o_Pk3D, o_PkX3D, o_phase, o_Nm1D, o_Pk1D, o_PkX1D, o_Nm2D, o_Pk2D, o_PkX2D;
};

__global__ void __launch_bounds__(256) pk_finalize_kernel(const FinalizeArgs A) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double *f = reinterpret_cast<double *>(A.acc);
    auto count = [&](long long w) -> double {
        return A.counts_f64 ? f[w] : (double)A.acc[w];
    };
    if (t < A.n3) {
        const double Nm = count(A.o_Nm3D + t);
        f[A.o_k3D + t] = __dmul_rn(__ddiv_rn(f[A.o_k3D + t], Nm), A.kF);
        const double ell[3] = {1.0, 5.0, 9.0};
        for (int l = 0; l < 3; l++) {
            for (int c = 0; c < A.F; c++) {
                const long long w = A.o_Pk3D + (t * 3 + l) * A.F + c;
                f[w] = __dmul_rn(__ddiv_rn(__dmul_rn(f[w], ell[l]), Nm), A.fact);
            }
            for (int c = 0; c < A.X; c++) {
                const long long w = A.o_PkX3D + (t * 3 + l) * A.X + c;
                f[w] = __dmul_rn(__ddiv_rn(__dmul_rn(f[w], ell[l]), Nm), A.fact);
            }
        }
        f[A.o_phase + t] = __dmul_rn(__ddiv_rn(f[A.o_phase + t], Nm), A.fact);
        f[A.o_Nm3D + t] = Nm;
    }
    if (t < A.n1) {
        const double Nm = count(A.o_Nm1D + t);
        const double k1 = __dmul_rn(__ddiv_rn(__dmul_rn(Nm, (double)t), Nm), A.kF);
        const double kmaxper = sqrt(__dsub_rn(A.kN2, __dmul_rn(k1, k1)));
        const double w1 = __ddiv_rn(__dmul_rn(M_PI, __dmul_rn(kmaxper, kmaxper)), Nm);
        const double twopi2 = A.twopi2;
        for (int c = 0; c < A.F; c++) {
            const long long w = A.o_Pk1D + t * A.F + c;
            f[w] = __ddiv_rn(__dmul_rn(__dmul_rn(f[w], A.fact), w1), twopi2);
        }
        for (int c = 0; c < A.X; c++) {
            const long long w = A.o_PkX1D + t * A.X + c;
            f[w] = __ddiv_rn(__dmul_rn(__dmul_rn(f[w], A.fact), w1), twopi2);
        }
        f[A.o_Nm1D + t] = Nm;
    }
    if (t < A.n2) {
        const double Nm = count(A.o_Nm2D + t);
        for (int c = 0; c < A.F; c++) {
            const long long w = A.o_Pk2D + t * A.F + c;
            f[w] = __ddiv_rn(__dmul_rn(f[w], A.fact), Nm);
        }
        for (int c = 0; c < A.X; c++) {
            const long long w = A.o_PkX2D + t * A.X + c;
            f[w] = __ddiv_rn(__dmul_rn(f[w], A.fact), Nm);
        }
        f[A.o_Nm2D + t] = Nm;
        if (A.kpar != nullptr) {
            const long long kp = t % A.kmax_par1, kq = t / A.kmax_par1;
            A.kpar[t] = __dmul_rn(0.5 * (double)(kp + kp + 1), A.kF);      // 0.5*(k_par + k_par+1)*kF
            A.kper[t] = __dmul_rn(0.5 * (double)(kq + kq + 1), A.kF);
        }
    }
}

// ---- spectra of delta = n/<n> - 1 straight from the transform of n (Pk_snapshot.py:88-89 fused away) -----------
// FFT(delta) = FFT(n)/<n> except for the DC mode, which the -1 cancels, and <n> = Re FFT(n)[0] / dims^3: taking
// the DC mode out of the spectrum and scaling the binned sums by 1/(<n>_i <n>_j) gives every spectrum of delta
// without the two passes over the grid (a float64 sum and the in-place n/<n> - 1).
// A float32 transform carries rounding noise proportional to its largest partial sums, and with the whole mass in
// the DC mode those sit on the three axes through k = 0 (relative amplitude error ~1e-7 dims/sigma there: 1e-4 at
// 512^3 of unclustered particles).  So the grid may hold n - c for any constant c near <n> (the deposit starts
// from -c instead of 0, see field.prebias_): then <n> = c + Re FFT[0]/dims^3 exactly, and the DC mode is small.
struct DcArgs {
    float *dk[PYL_MAX_FIELDS];
    int F, owner;
    double *dc;
};

__global__ void pk_take_dc_kernel(const DcArgs A) {
    const int f = threadIdx.x;
    if (f >= A.F) return;
    double v = 0.0;
    if (A.owner) {
        v = (double)A.dk[f][0];
        A.dk[f][0] = 0.0f;
        A.dk[f][1] = 0.0f;
    }
    A.dc[f] = v;
}

struct DensityScaleArgs {
    double *f;
    const double *dc, *offset;
    double cells;
    int F, X;
    long long o[6], n[6];          // Pk3D, PkX3D, Pk1D, PkX1D, Pk2D, PkX2D: first word, words
};

__global__ void __launch_bounds__(256) pk_density_scale_kernel(const DensityScaleArgs A) {
    __shared__ double s_auto[PYL_MAX_FIELDS], s_cross[PYL_MAX_FIELDS * (PYL_MAX_FIELDS - 1) / 2 + 1];
    if (threadIdx.x == 0) {
        double inv[PYL_MAX_FIELDS];
        for (int c = 0; c < A.F; c++)                                      // 1 / <n>_c
            inv[c] = 1.0 / ((A.offset ? A.offset[c] : 0.0) + A.dc[c] / A.cells);
        int x = 0;
        for (int i = 0; i < A.F; i++) {
            s_auto[i] = inv[i] * inv[i];
            for (int j = i + 1; j < A.F; j++) s_cross[x++] = inv[i] * inv[j];
        }
    }
    __syncthreads();
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int r = 0; r < 6; r++) {
        if (t < A.n[r]) {
            const bool cross = r & 1;
            const int per = cross ? A.X : A.F;
            const int c = (int)(t % per);
            A.f[A.o[r] + t] *= cross ? s_cross[c] : s_auto[c];
            return;
        }
        t -= A.n[r];
    }
}

static void fill_layout(int dims, int F, pyl_pk_layout_t *L) {
    const int m = dims / 2;
    const int X = F * (F - 1) / 2;
    L->dims = dims; L->fields = F; L->xfields = X;
    L->kmax_par = m;
    L->kmax_per = (int)sqrt((double)m * m + (double)m * m);
    L->kmax = (int)sqrt(3.0 * (double)m * m);
    L->n2d = (int64_t)(L->kmax_par + 1) * (L->kmax_per + 1);
    const int64_t n3 = L->kmax + 1, n1 = L->kmax_par + 1;
    int64_t w = 0;
    L->k3D = w; w += n3;
    L->Nm3D = w; w += n3;
    L->Pk3D = w; w += n3 * 3 * F;
    L->PkX3D = w; w += n3 * 3 * X;
    L->phase = w; w += n3;
    L->Nm1D = w; w += n1;
    L->Pk1D = w; w += n1 * F;
    L->PkX1D = w; w += n1 * X;
    L->Nm2D = w; w += L->n2d;
    L->Pk2D = w; w += L->n2d * F;
    L->PkX2D = w; w += L->n2d * X;
    L->total_words = w;
}

static size_t bin_workspace(int dims, int F) {
    pyl_pk_layout_t L;
    fill_layout(dims, F, &L);
    const size_t tab = align_up((size_t)F * (dims / 2 + 1) * sizeof(double), 256);
    const size_t rep = (size_t)PK_NREP * (size_t)L.Nm2D * 8;
    return tab + rep;
}

// rows N-ky of the lower rows ky in [ky_lo, ky_lo+ny_lo) that exist as separate modes (ky != 0, ky != Nyquist),
// stored in ascending order of their global index; *first = global index of the first of them
static int mirrored_upper_rows(int dims, int ky_lo, int ny_lo, int *first) {
    const int m = dims / 2;
    int lo = ky_lo, hi = ky_lo + ny_lo - 1;             // |ky| values with a mirror: lo..hi minus {0, Nyquist}
    if (lo == 0) lo = 1;
    if (dims % 2 == 0 && hi == m) hi = m - 1;
    if (first) *first = dims - hi;
    return hi >= lo ? hi - lo + 1 : 0;
}

// ---- per-geometry cache of the field-independent sections ------------------------------------------------
// Nmodes3D/1D/2D and the per-shell sum of |k| depend on (dims, axis, which ky rows the caller holds) only.  The
// first call for a geometry accumulates them (GEOM kernel) and keeps a device copy; later calls run the lighter
// kernel and copy the four sections back in -- like a cuFFT plan, the copy is made once per geometry and lives
// until pyl_pk_clear_cache().
struct GeomCache {
    unsigned long long *data = nullptr;     // [k3D n3 | Nm3D n3 | Nm1D n1 | Nm2D n2]
};
using GeomKey = std::tuple<int, int, int, int, int, int>;   // device, dims, ky_lo, nky, mirrored, axis
static std::map<GeomKey, GeomCache> g_geom;
static std::mutex g_geom_mu;

template <int F>
static int launch_bin(const float *const *delta_k, const int *mas_index, int dims, int ky_lo,
                      int nky, int mirrored, int axis, int flags, void *out, void *ws, cudaStream_t stream) {
    const int want_phase = flags & PYL_PK_PHASE;
    pyl_pk_layout_t L;
    fill_layout(dims, F, &L);
    const int m = dims / 2, nz = m + 1;

    double *tab = reinterpret_cast<double *>(ws);
    const size_t tab_bytes = align_up((size_t)F * nz * sizeof(double), 256);
    unsigned long long *rep =
        reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(ws) + tab_bytes);
    const long long rep_words = L.Nm2D;   // the 3D and 1D sections precede the 2D section

    PYL_CUDA_CHECK(cudaMemsetAsync(out, 0, (size_t)L.total_words * 8, stream));
    PYL_CUDA_CHECK(cudaMemsetAsync(rep, 0, (size_t)PK_NREP * rep_words * 8, stream));

    int p[4] = {0, 0, 0, 0};
    for (int f = 0; f < F; f++) p[f] = mas_index[f];
    pk_window_kernel<<<(nz + 127) / 128, 128, 0, stream>>>(tab, F, nz, dims, make_int4(p[0], p[1], p[2], p[3]));
    PYL_LAUNCH_CHECK();

    PkArgs<F> A;
    for (int f = 0; f < F; f++) {
        A.dk[f] = reinterpret_cast<const float2 *>(delta_k[f]);
        A.win[f] = tab + (size_t)f * nz;
    }
    A.N = dims; A.m = m; A.nz = nz; A.even = (dims % 2 == 0);
    A.ky_lo = ky_lo; A.nky = nky; A.ny_lo = nky; A.ky_up0 = 0;
    A.o_lo = 0; A.w_lo = 0; A.w_hi = m + 1;
    const bool whole = (ky_lo == 0 && nky == dims && !mirrored);
    if (whole) {
        // fold both in-plane axes; walk along y unless y is the line of sight
        A.fold_other = 1;
        A.walk_y = (axis == 1) ? 0 : 1;
        A.n_other = m + 1;
    } else if (mirrored) {
        // slab holding |ky| in [ky_lo, ky_lo+nky) AND the mirrored rows N-ky (pyl_pk_bin_mirrored; here `nky`
        // counts the lower rows): the whole-grid scheme restricted to a |ky| window -- four-fold mode sharing,
        // and the walk still avoids the line of sight (walk y over the window, or x when y is the line of sight)
        A.ny_lo = nky;
        A.nky = nky + mirrored_upper_rows(dims, ky_lo, nky, &A.ky_up0);
        A.fold_other = 1;
        if (axis == 1) { A.walk_y = 0; A.n_other = nky; A.o_lo = ky_lo; }
        else { A.walk_y = 1; A.n_other = m + 1; A.w_lo = ky_lo; A.w_hi = ky_lo + nky; }
    } else {
        // contiguous window of ky rows without their mirrors: only kx is complete -> walk x, no ky fold
        A.fold_other = 0;
        A.walk_y = 0;
        A.n_other = nky;
    }
    A.axis = axis;
    A.cross_imag = (flags & PYL_PK_CROSS_IMAG) ? 1 : 0;
    A.ky_major = (flags & PYL_PK_KY_MAJOR) ? 1 : 0;
    A.kmax_par1 = L.kmax_par + 1;
    A.nzp = (nz + 31) / 32 * 32;
    A.T = (long long)A.n_other * A.nzp;

    // cut the walk so that the grid has a few waves of warps even for small grids, and so that a warp's shell
    // window (PK_WBINS) covers its segment
    const long long warps_per_seg = (A.T + 31) / 32;
    const long long want_warps = (long long)sm_count() * 64;
    long long nseg = (want_warps + warps_per_seg - 1) / warps_per_seg;
    const int wlen = A.w_hi - A.w_lo;
    const long long max_seg = (wlen + 7) / 8;           // at least 8 steps per segment
    if (nseg > max_seg) nseg = max_seg;
    const long long min_seg = (wlen + PK_MAX_SEG - 1) / PK_MAX_SEG;
    if (nseg < min_seg) nseg = min_seg;
    if (nseg < 1) nseg = 1;
    A.seg_len = (int)((wlen + nseg - 1) / nseg);
    if (A.seg_len < 1) A.seg_len = 1;
    A.nseg = (wlen + A.seg_len - 1) / A.seg_len;

    A.out = reinterpret_cast<unsigned long long *>(out);
    A.rep = rep; A.rep_words = rep_words;
    A.o_k3D = L.k3D; A.o_Nm3D = L.Nm3D; A.o_Pk3D = L.Pk3D; A.o_PkX3D = L.PkX3D; A.o_phase = L.phase;
    A.o_Nm1D = L.Nm1D; A.o_Pk1D = L.Pk1D; A.o_PkX1D = L.PkX1D;
    A.o_Nm2D = L.Nm2D; A.o_Pk2D = L.Pk2D; A.o_PkX2D = L.PkX2D;

    int dev = 0;
    PYL_CUDA_CHECK(cudaGetDevice(&dev));
    const GeomKey key(dev, dims, ky_lo, nky, mirrored, axis);
    const long long n3 = L.kmax + 1, n1 = L.kmax_par + 1, n2 = L.n2d;
    unsigned long long *cached = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_geom_mu);
        auto it = g_geom.find(key);
        if (it != g_geom.end()) cached = it->second.data;
    }

    if (A.T > 0 && A.nseg > 0) {
        dim3 grid((unsigned)((A.T + PK_BLOCK - 1) / PK_BLOCK), (unsigned)A.nseg);
        constexpr int X = F * (F - 1) / 2;
        const size_t smem = (size_t)(PK_BLOCK / 32) * PK_WBINS * (3 * F + 3 * X + (want_phase ? 1 : 0)) * 8;
        auto launch = [&](auto kernel) -> int {
            if (smem > 48 * 1024)
                PYL_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kernel<<<grid, PK_BLOCK, smem, stream>>>(A);
            PYL_LAUNCH_CHECK();
            return PYL_OK;
        };
        int st;
        if (cached) st = want_phase ? launch(pk_bin_walk_kernel<F, true, false>) : launch(pk_bin_walk_kernel<F, false, false>);
        else st = want_phase ? launch(pk_bin_walk_kernel<F, true, true>) : launch(pk_bin_walk_kernel<F, false, true>);
        if (st != PYL_OK) return st;
    }
    pk_fold_replicas_kernel<<<(unsigned)((rep_words + 255) / 256), 256, 0, stream>>>(
        A.out, rep, rep_words, PK_NREP, L.Nm3D, (long long)L.kmax + 1, L.Nm1D, (long long)L.kmax_par + 1);
    PYL_LAUNCH_CHECK();

    unsigned long long *o = A.out;
    const cudaMemcpyKind d2d = cudaMemcpyDeviceToDevice;
    if (cached) {
        PYL_CUDA_CHECK(cudaMemcpyAsync(o + L.k3D, cached, (size_t)n3 * 8, d2d, stream));
        PYL_CUDA_CHECK(cudaMemcpyAsync(o + L.Nm3D, cached + n3, (size_t)n3 * 8, d2d, stream));
        PYL_CUDA_CHECK(cudaMemcpyAsync(o + L.Nm1D, cached + 2 * n3, (size_t)n1 * 8, d2d, stream));
        PYL_CUDA_CHECK(cudaMemcpyAsync(o + L.Nm2D, cached + 2 * n3 + n1, (size_t)n2 * 8, d2d, stream));
    } else {
        unsigned long long *c = nullptr;
        PYL_CUDA_CHECK(cudaMalloc(&c, (size_t)(2 * n3 + n1 + n2) * 8));
        PYL_CUDA_CHECK(cudaMemcpyAsync(c, o + L.k3D, (size_t)n3 * 8, d2d, stream));
        PYL_CUDA_CHECK(cudaMemcpyAsync(c + n3, o + L.Nm3D, (size_t)n3 * 8, d2d, stream));
        PYL_CUDA_CHECK(cudaMemcpyAsync(c + 2 * n3, o + L.Nm1D, (size_t)n1 * 8, d2d, stream));
        PYL_CUDA_CHECK(cudaMemcpyAsync(c + 2 * n3 + n1, o + L.Nm2D, (size_t)n2 * 8, d2d, stream));
        PYL_CUDA_CHECK(cudaStreamSynchronize(stream));      // once per geometry: the copy is complete before it is published
        std::lock_guard<std::mutex> lock(g_geom_mu);
        if (g_geom.find(key) == g_geom.end()) g_geom[key].data = c;
        else cudaFree(c);
    }
    return PYL_OK;
}

}  // namespace pyl

using namespace pyl;

extern "C" {

int pyl_pk_layout(int dims, int fields, pyl_pk_layout_t *layout) {
    PYL_REQUIRE(layout != nullptr, "pyl_pk_layout: layout is NULL");
    PYL_REQUIRE(dims > 0 && fields >= 1, "pyl_pk_layout: bad dims/fields");
    fill_layout(dims, fields, layout);
    return PYL_OK;
}

int pyl_pk_clear_cache(void) {
    int dev = 0;
    PYL_CUDA_CHECK(cudaGetDevice(&dev));
    PYL_CUDA_CHECK(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lock(g_geom_mu);
    for (auto it = g_geom.begin(); it != g_geom.end();) {
        if (std::get<0>(it->first) == dev) {
            cudaFree(it->second.data);
            it = g_geom.erase(it);
        } else {
            ++it;
        }
    }
    return PYL_OK;
}

int pyl_pk_counts_to_f64(void *acc, int dims, int fields, pyl_stream_t stream) {
    PYL_REQUIRE(acc != nullptr, "pyl_pk_counts_to_f64: NULL buffer");
    PYL_REQUIRE(dims > 0 && fields >= 1, "pyl_pk_counts_to_f64: bad dims/fields");
    pyl_pk_layout_t L;
    fill_layout(dims, fields, &L);
    const long long n0 = L.kmax + 1, n1 = L.kmax_par + 1, n2 = L.n2d;
    const long long n = n0 + n1 + n2;
    pk_counts_to_f64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<unsigned long long *>(acc), L.Nm3D, n0, L.Nm1D, n1, L.Nm2D, n2);
    PYL_LAUNCH_CHECK();
    return PYL_OK;
}

int pyl_pk_finalize(void *acc, int dims, int fields, double BoxSize, int counts_are_f64, double *kpar,
                    double *kper, pyl_stream_t stream) {
    PYL_REQUIRE(acc != nullptr, "pyl_pk_finalize: NULL buffer");
    PYL_REQUIRE((kpar == nullptr) == (kper == nullptr), "pyl_pk_finalize: give both kpar and kper or neither");
    PYL_REQUIRE(dims > 0 && fields >= 1 && BoxSize > 0.0, "pyl_pk_finalize: bad dims/fields/BoxSize");
    pyl_pk_layout_t L;
    fill_layout(dims, fields, &L);
    FinalizeArgs A;
    A.acc = reinterpret_cast<unsigned long long *>(acc);
    A.kpar = kpar; A.kper = kper; A.kmax_par1 = L.kmax_par + 1;
    A.F = L.fields; A.X = L.xfields;
    A.n3 = L.kmax + 1; A.n1 = L.kmax_par + 1; A.n2 = L.n2d;
    A.counts_f64 = counts_are_f64;
    A.kF = 2.0 * M_PI / BoxSize;                                   // Pk_library.pyx:57
    A.kN2 = pow((double)(dims / 2) * A.kF, 2.0);                   // kN**2, :59 and :389
    A.fact = pow(BoxSize / ((double)dims * (double)dims), 3.0);    // (BoxSize/dims**2)**3, :385
    A.twopi2 = pow(2.0 * M_PI, 2.0);                               // (2.0*np.pi)**2, :391
    A.o_k3D = L.k3D; A.o_Nm3D = L.Nm3D; A.o_Pk3D = L.Pk3D; A.o_PkX3D = L.PkX3D; A.o_phase = L.phase;
    A.o_Nm1D = L.Nm1D; A.o_Pk1D = L.Pk1D; A.o_PkX1D = L.PkX1D;
    A.o_Nm2D = L.Nm2D; A.o_Pk2D = L.Pk2D; A.o_PkX2D = L.PkX2D;
    long long n = A.n2 > A.n3 ? A.n2 : A.n3;
    if (A.n1 > n) n = A.n1;
    pk_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(A);
    PYL_LAUNCH_CHECK();
    return PYL_OK;
}

int pyl_pk_take_dc(float *const *delta_k, int fields, int holds_dc, double *dc, pyl_stream_t stream) {
    PYL_REQUIRE(fields >= 1 && fields <= PYL_MAX_FIELDS, "pyl_pk_take_dc: fields must be 1..PYL_MAX_FIELDS");
    PYL_REQUIRE(delta_k != nullptr && dc != nullptr, "pyl_pk_take_dc: NULL pointer");
    DcArgs A;
    A.F = fields; A.owner = holds_dc ? 1 : 0; A.dc = dc;
    for (int f = 0; f < fields; f++) {
        PYL_REQUIRE(delta_k[f] != nullptr || !holds_dc, "pyl_pk_take_dc: NULL field pointer");
        A.dk[f] = delta_k[f];
    }
    pk_take_dc_kernel<<<1, 32, 0, as_stream(stream)>>>(A);
    PYL_LAUNCH_CHECK();
    return PYL_OK;
}

int pyl_pk_density_scale(void *acc, int dims, int fields, const double *dc, const double *offset,
                         pyl_stream_t stream) {
    PYL_REQUIRE(acc != nullptr && dc != nullptr, "pyl_pk_density_scale: NULL pointer");
    PYL_REQUIRE(dims > 0 && fields >= 1 && fields <= PYL_MAX_FIELDS, "pyl_pk_density_scale: bad dims/fields");
    pyl_pk_layout_t L;
    fill_layout(dims, fields, &L);
    DensityScaleArgs A;
    A.f = reinterpret_cast<double *>(acc);
    A.dc = dc;
    A.offset = offset;
    A.cells = (double)dims * (double)dims * (double)dims;
    A.F = L.fields; A.X = L.xfields;
    const long long n3 = L.kmax + 1, n1 = L.kmax_par + 1;
    const long long off[6] = {L.Pk3D, L.PkX3D, L.Pk1D, L.PkX1D, L.Pk2D, L.PkX2D};
    const long long cnt[6] = {n3 * 3 * A.F, n3 * 3 * A.X, n1 * A.F, n1 * A.X, L.n2d * A.F, L.n2d * A.X};
    long long total = 0;
    for (int r = 0; r < 6; r++) { A.o[r] = off[r]; A.n[r] = cnt[r]; total += cnt[r]; }
    pk_density_scale_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(A);
    PYL_LAUNCH_CHECK();
    return PYL_OK;
}

size_t pyl_pk_bin_workspace_bytes(int dims, int fields) {
    if (dims <= 0 || fields < 1 || fields > PYL_MAX_FIELDS) return 0;
    return bin_workspace(dims, fields);
}

static int pk_bin_entry(const float *const *delta_k, int fields, const int *mas_index, int dims,
                        int ky_lo, int nky, int mirrored, int axis, int flags, void *out, void *ws,
                        size_t ws_bytes, pyl_stream_t stream) {
    const int want_phase = flags & PYL_PK_PHASE;
    PYL_REQUIRE(fields >= 1 && fields <= PYL_MAX_FIELDS, "pyl_pk_bin: fields must be 1..PYL_MAX_FIELDS");
    PYL_REQUIRE(dims > 0 && dims <= 8192, "pyl_pk_bin: dims must be in 1..8192");
    PYL_REQUIRE(axis >= 0 && axis <= 2, "pyl_pk_bin: axis must be 0, 1 or 2");
    PYL_REQUIRE(delta_k != nullptr && mas_index != nullptr && out != nullptr, "pyl_pk_bin: NULL pointer");
    PYL_REQUIRE(ky_lo >= 0 && nky >= 0 && ky_lo + nky <= (mirrored ? dims / 2 + 1 : dims), "pyl_pk_bin: bad ky window");
    for (int f = 0; f < fields; f++) {
        PYL_REQUIRE(delta_k[f] != nullptr || nky == 0, "pyl_pk_bin: NULL field pointer");
        PYL_REQUIRE(mas_index[f] >= 0 && mas_index[f] <= 4, "pyl_pk_bin: mas_index must be 0..4");
    }
    PYL_REQUIRE(!want_phase || fields == 1, "pyl_pk_bin: phase is accumulated for a single field only");
    const size_t need = bin_workspace(dims, fields);
    if (ws == nullptr || ws_bytes < need) {
        set_last_error("pyl_pk_bin: workspace of %zu bytes required, %zu given", need, ws_bytes);
        return PYL_ERR_WORKSPACE;
    }
    cudaStream_t s = as_stream(stream);
    switch (fields) {
        case 1: return launch_bin<1>(delta_k, mas_index, dims, ky_lo, nky, mirrored, axis, flags, out, ws, s);
        case 2: return launch_bin<2>(delta_k, mas_index, dims, ky_lo, nky, mirrored, axis, flags & ~PYL_PK_PHASE, out, ws, s);
        case 3: return launch_bin<3>(delta_k, mas_index, dims, ky_lo, nky, mirrored, axis, flags & ~PYL_PK_PHASE, out, ws, s);
        default: return launch_bin<4>(delta_k, mas_index, dims, ky_lo, nky, mirrored, axis, flags & ~PYL_PK_PHASE, out, ws, s);
    }
}

int pyl_pk_bin(const float *const *delta_k, int fields, const int *mas_index, int dims,
               int ky_lo, int nky, int axis, int want_phase, void *out, void *ws,
               size_t ws_bytes, pyl_stream_t stream) {
    return pk_bin_entry(delta_k, fields, mas_index, dims, ky_lo, nky, 0, axis, want_phase, out, ws, ws_bytes, stream);
}

int pyl_pk_mirrored_rows(int dims, int ky_lo, int ny_lo, int *first_upper) {
    if (dims <= 0 || ky_lo < 0 || ny_lo < 0 || ky_lo + ny_lo > dims / 2 + 1) return -1;
    return ny_lo + mirrored_upper_rows(dims, ky_lo, ny_lo, first_upper);
}

int pyl_pk_bin_mirrored(const float *const *delta_k, int fields, const int *mas_index, int dims,
                        int ky_lo, int ny_lo, int axis, int want_phase, void *out, void *ws,
                        size_t ws_bytes, pyl_stream_t stream) {
    return pk_bin_entry(delta_k, fields, mas_index, dims, ky_lo, ny_lo, 1, axis, want_phase, out, ws, ws_bytes, stream);
}

}  // extern "C"
