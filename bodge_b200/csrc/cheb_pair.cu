// Two Chebyshev steps per pass over the vectors ("pair step", SURVEY 8f-4: temporal blocking).
//
//     T_{n+1} = 2 H~ T_n - T_{n-1},     T_{n+2} = 2 H~ T_{n+1} - T_n
//
// in ONE launch, on the block-dictionary matrix format of cheb_ell.cu.  A single-step kernel moves
// three vector passes per step (read T_n, read T_{n-1}, write T_{n+1}); this one moves four per TWO
// steps (read T_{n-1}, T_n; write T_{n+1}, T_{n+2}) because T_{n+1} never has to come back from HBM:
// it is consumed out of shared memory.  With the matrix already out of the HBM stream (dictionary
// format) that is the only way past the one-pass roofline of the step.
//
// Geometry.  Lattices whose x-planes are one-dimensional (Lz = 1 or Ly = 1; M >= 3 sites per plane, Lx >= 3 planes),
// treated as a TORUS: open and periodic stencils run the same code, the matrix (direction codes) decides what is
// connected.  The plane is cut into patches of P owned sites; a CTA marches one patch along a segment of x:
//
//   iteration i:  [A] T_{n+1} of plane x0 - 1 + i for the P owned sites AND one halo site on either side, from three
//                     T_n planes (P + 4 sites each) staged in a shared-memory ring of 8 planes by bulk async copies
//                     (TMA: cp.async.bulk, one contiguous 9 KB run per plane -- plus one small run from the opposite
//                     side of the plane for the first / last patch --, seven planes ahead, completion on an mbarrier),
//                     and T_{n-1} of the warp's rows, which only this warp reads: straight to registers, one plane
//                     ahead; the result goes to a shared-memory ring (4 planes) and, for owned sites of owned planes,
//                     to HBM;
//                 __syncthreads; the plane [A] no longer needs is replaced by the plane eight ahead
//                 [B] T_{n+2} of the plane one behind, for the owned sites, from the T_{n+1} ring's three planes -- of which
//                     only the middle one, written an iteration ago, is read at other warps' sites -- and the T_n records
//                     [A] read as its x-1 neighbours (kept in registers across the barrier).
//
// The halo T_{n+1} values (one site either side in y, one plane either side of the segment in x)
// are recomputed, not exchanged: (P+2)/P x (len+2)/len redundant work on sub-step [A], no
// inter-CTA dependency, deterministic.  Because a neighbouring CTA may still need T_{n-1} / T_n of
// a site after its owner has produced T_{n+1} / T_{n+2} there, the outputs go to two further
// buffers (four vector buffers in rotation) instead of in place.
//
// Arithmetic per row is exactly that of cheb_step_ell<.., DICT, DIAG> (same fragments, same order,
// same update expression): on open lattices the vectors are bit-identical to the single-step dictionary kernels'
// (where a neighbour wraps around, the terms of a row are added in stencil direction here and in ascending block
// column there: equal to rounding); the dot products are summed over a different partition of the rows (agree to
// rounding).
//
// Compile-time experiments kept for the record (all within +-1 % of the default at 10^6 sites):
// -DBDG_PAIR_PV3 (E_{j-1} two planes ahead), -DBDG_PAIR_SPLIT (split-phase mbarrier hand-over between [A] and [B],
// [B] started on the warp's own records before the wait); run time: BDG_PAIR_WARPS=12 (REG: own records in registers,
// one CTA per SM), =16; BDG_PAIR_SELF, BDG_PAIR_P, BDG_PAIR_SEG.
#include <algorithm>
#include <array>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "bdg_internal.h"
#include "cheb_device.cuh"
#include "cheb_smem.cuh"
#include "work_lists.h"

namespace {

constexpr int kRecBytes = 512;  // one site record at PW = 8: 8 columns x 4 components x complex128
constexpr int kRing = 4;        // planes of the T_{n+1} ring (three read by [B], one being written)
constexpr int kRingN = 8;       // planes of the T_n ring: three in use, five in flight (power of two)
// Hand-over between the warps of a CTA: the __syncthreads between [A] and [B] (default); -DBDG_PAIR_LAZY = arrive at the end of an
// iteration, wait after [A] of the next (see the loop: +2.8 % at burst clocks, -1 % under the sustained power cap: its 256
// polling threads cost the clock what the slack gains; one poller per warp is 9 % slower; profiles/r02/54_* .. 58_*);
// -DBDG_PAIR_SPLIT = the earlier split-phase experiment.
#ifdef BDG_PAIR_LAZY
constexpr int kRingGuard = 1;
#else
constexpr int kRingGuard = 0;
#endif

constexpr int kDirs = 5;  // direction-ordered row: self, x-1, y-1, y+1, x+1 (= ascending block column)

// B fragments of the S rows a warp handles (cheb_ell.cu: DICT / DIAG), held in registers from plane to
// plane and reloaded only when a code differs from the held one.  Lane s * 5 + u carries the code of
// row s, direction u; code < 0 = no block.  SELF: the on-site fragments (u = 0) are not held -- they are
// fetched per row, one plane ahead (self_fragments) -- so their codes do not take part in the test.
template <bool DIAG, bool SELF, int S, bool SD = false>
__device__ __forceinline__ void hold_fragments(int jv, int &jheld, double (&keep)[S][kDirs], const double *__restrict__ table,
                                               const double *__restrict__ dtab, int lane, bool self_lane) {
    const unsigned changed = __ballot_sync(kFull, jv != jheld && !(SELF && self_lane));
    if (changed) {
        jheld = jv;
#pragma unroll
        for (int s = 0; s < S; ++s) {
#pragma unroll
            for (int u = SELF ? 1 : 0; u < kDirs; ++u) {
                const int code = __shfl_sync(kFull, jv, s * kDirs + u);
                const unsigned take = changed >> (s * kDirs + u) & 1u;
                const size_t c = (size_t)max(code, 0);
                const double *entry = ((DIAG && u > 0) || (SD && u == 0)) ? dtab + c * 4 + (lane & 3) : table + c * 32 + lane;
                ld_table_pred(keep[s][u], entry, take && code >= 0);
                if (take && code < 0) keep[s][u] = 0.0;
            }
        }
    }
}

// SELF: this lane's double of the on-site B fragment of row s, straight from the table by the row's code
// (-1: no diagonal block -> 0).  Issued a plane ahead of its use; with every on-site block distinct (disorder,
// self-consistent gap) this is the 256-byte-per-site stream the matrix cannot do without, with a phase winding
// or a layered structure it hits in L1 / L2.
template <int S, bool SD = false>
__device__ __forceinline__ void self_fragments(int jv, double (&f)[S], const double *__restrict__ table, const double *__restrict__ dtab,
                                               int lane) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int code = __shfl_sync(kFull, jv, s * kDirs);
        f[s] = 0.0;
        ld_table_pred(f[s], SD ? dtab + (size_t)max(code, 0) * 4 + (lane & 3) : table + (size_t)max(code, 0) * 32 + lane, code >= 0);
    }
}

// y = sum_u B_u x_u for this lane's element; order and operations of cheb_step_ell (directions: self, x-1, y-1,
// y+1, x+1 = ascending block column on an open lattice), in stages so that sub-step [B] can start on the records its
// own warp wrote while the rest of the CTA is still arriving at the barrier: begin (self), add x 4, end.
// SD (without DIAG): the on-site block is real and diagonal (cheb_ell.cu: SD) -- its product is two multiplications, and
// the accumulators of the hopping blocks' MMAs start from them (what the MMA of such a block leaves there).
template <bool DIAG, bool SD = false> struct RowSum {
    double a10 = 0.0, a11 = 0.0, a20 = 0.0, a21 = 0.0, yr = 0.0, yi = 0.0;
    __device__ __forceinline__ void begin(const double2 &x0, double b0) {
        if (SD && !DIAG) {
            a10 = b0 * x0.x, a20 = b0 * x0.y;
            return;
        }
        dmma_8x8x4(a10, a11, x0.x, b0);
        dmma_8x8x4(a20, a21, x0.y, b0);
        if (DIAG) yr = a10 - a21, yi = a11 + a20;
    }
    __device__ __forceinline__ void add(const double2 &x, double b) {
        if (DIAG) {
            yr = fma(b, x.x, yr);
            yi = fma(b, x.y, yi);
        } else {
            dmma_8x8x4(a10, a11, x.x, b);
            dmma_8x8x4(a20, a21, x.y, b);
        }
    }
    __device__ __forceinline__ void end() {
        if (!DIAG) yr = a10 - a21, yi = a11 + a20;
    }
};

template <bool DIAG, bool SD = false>
__device__ __forceinline__ void row_product(const double2 &x0, const double2 &x1, const double2 &x2, const double2 &x3,
                                            const double2 &x4, double b0, const double (&bop)[kDirs], double &yr, double &yi) {
    RowSum<DIAG, SD> r;
    r.begin(x0, b0);
    r.add(x1, bop[1]);
    r.add(x2, bop[2]);
    r.add(x3, bop[3]);
    r.add(x4, bop[4]);
    r.end();
    yr = r.yr;
    yi = r.yi;
}

#if defined(BDG_PAIR_SPLIT) || defined(BDG_PAIR_LAZY)
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
#endif

// NW warps per CTA, two ADJACENT sites per warp and plane: W = 2 NW sites per plane in sub-step [A]
// (P <= W - 2 of them owned).  Dynamic shared memory: kRingN planes of W + 2 records (T_n), kRing
// planes of W records (T_{n+1}), two guard records, one mbarrier per T_n plane: 105 KB at NW = 8, so two
// CTAs share an SM (one computes while the other waits at its barrier).
//
// Geometry is a TORUS: planes x and in-plane sites y are taken modulo Lx and M, so the halo of the first /
// last patch and of the first / last segment is the opposite face of the lattice (staged by its own small bulk
// copy).  What is connected is decided by the matrix alone: it arrives as `dcode[row][5]`, the dictionary code
// of the row's block in each stencil direction (self, x-1, y-1, y+1, x+1; -1 = none), so the reference's
// periodic skeleton with the edges filled in (bodge/lattice.py:161-197) and the open lattice it leaves when they
// are not run the same code: a direction without a block meets a zero fragment (whatever finite vector value
// sits at the wrapped position is multiplied by zero; the rings are cleared once).  The records a row needs sit
// at fixed offsets from its own in the rings -- no per-row index arithmetic -- and the two sites of a warp
// share their in-plane neighbours (4 loads for 6 operands).  Warps whose site does not exist (ragged last patch)
// compute on whatever the rings hold and store nothing: the row body has no branches.
//
// An item = (patch, segment [x0, x0 + len)).  Iteration i = 0 .. len + 1 of an item:
//   [A] plane x0 - 1 + i from the T_n planes t = i, i + 1, i + 2 of the item (t <-> plane x0 - 2 + t);
//       stored / counted in the dot products for 1 <= i <= len
//   [B] plane x0 - 2 + i (i >= 2) from the T_{n+1} planes of iterations i - 2, i - 1, i
// T_n planes occupy ring slots in the order they are consumed, across items (`cnt`): slot = count & 7, mbarrier
// parity = (count >> 3) & 1.
//
// MODE 1 ("T2", the doubled-argument recursion) runs the same two sub-steps on the EVEN vectors only:
//     E_j = T_2j(H~) x,   E_{j+1} = 2 T_2(H~) E_j - E_{j-1} = 4 H~ (H~ E_j) - 2 E_j - E_{j-1}
// [A] u = H~ E_j (halo included, shared memory only, never stored), [B] E_{j+1} for the owned rows, written in
// place over E_{j-1} (read by its owner alone).  One launch is still two applications of H~ and four dot products
//     a = <E_j, E_j>, c = <u, E_j>, b = <E_{j+1}, E_j>, d = <E_{j+1}, u>
// from which the same four moments follow (cheb.cu: t2_normalize), but it moves THREE vector passes instead
// of four, and two vector buffers suffice.  `first`: E_1 = T_2(H~) E_0 = 2 H~ u - E_0 (E_{-1} = E_1).
//
// REG (the 12-warp shape: one CTA per SM, 168 registers per thread): the warp's OWN records stay in registers from
// plane to plane -- the T_n records it read as its x+1 neighbours are its centre records one iteration later and its x-1
// neighbours (and the T_n of [B]) two iterations later, and the T_{n+1} records it computes are its own operands of [B]
// for three iterations -- so shared memory is read only for what OTHER warps own: 6 instead of 16 LDS.128 per warp and
// iteration.  The shared-memory / L1 data pipe is what bounds the 8-warp shape.
// The four dot products of a run (= the pieces of one panel a CTA works through): quad -> warp -> CTA -> partials[run]; the
// last run of a panel to arrive (runs r0 .. r1) adds the panel's partials up in a fixed order.  Called by all threads between
// two pieces or at the end: the rings are idle and serve as scratch.
template <int NW>
__device__ __forceinline__ void flush_dots(double d0, double d1, double d2, double d3, unsigned char *scratch, int run, int r0, int r1,
                                           int panel, int n_panels, double *__restrict__ partials, unsigned *__restrict__ tickets,
                                           double *__restrict__ dots_step) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ bool is_last;
    d0 += __shfl_xor_sync(kFull, d0, 1);
    d1 += __shfl_xor_sync(kFull, d1, 1);
    d2 += __shfl_xor_sync(kFull, d2, 1);
    d3 += __shfl_xor_sync(kFull, d3, 1);
    d0 += __shfl_xor_sync(kFull, d0, 2);
    d1 += __shfl_xor_sync(kFull, d1, 2);
    d2 += __shfl_xor_sync(kFull, d2, 2);
    d3 += __shfl_xor_sync(kFull, d3, 2);
    __syncthreads();
    double *red = reinterpret_cast<double *>(scratch);  // [NW][4][8], then comb [NW][32] behind it
    double *comb = red + NW * 32;
    if ((lane & 3) == 0) {
        const int col = lane >> 2;
        red[(warp * 4 + 0) * 8 + col] = d0;
        red[(warp * 4 + 1) * 8 + col] = d1;
        red[(warp * 4 + 2) * 8 + col] = d2;
        red[(warp * 4 + 3) * 8 + col] = d3;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += red[w * 32 + threadIdx.x];
        partials[(size_t)run * 32 + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(&tickets[panel], 1u) == (unsigned)(r1 - r0 - 1);
    __syncthreads();
    if (is_last) {
        __threadfence();
        double s = 0.0;
        for (int b = r0 + warp; b < r1; b += NW) s += __ldcg(&partials[(size_t)b * 32 + lane]);
        comb[warp * 32 + lane] = s;
        __syncthreads();
        if (threadIdx.x < 32) {
            double t = 0.0;
#pragma unroll
            for (int g = 0; g < NW; ++g) t += comb[g * 32 + threadIdx.x];
            // slot = which * 8 + column; which = (step within the pair) * 2 + (0: <T,T>, 1: <T',T>)
            dots_step[(size_t)(threadIdx.x >> 3) * n_panels * 8 + panel * 8 + (threadIdx.x & 7)] = t;
        }
        if (threadIdx.x == 0) tickets[panel] = 0u;
    }
    __syncthreads();  // scratch and is_last are free again
}

// LISTED: the CTA's pieces come from the work lists of PairWalk (balanced plan, grid = (n_ctas, 1)); otherwise they are
// computed from the item index (classic plan, grid = (CTAs per panel, panels)) -- kept as arithmetic on kernel parameters
// because values loaded from memory leave the uniform datapath and the plane loop's predicates with them (+2 % time).
template <bool DIAG, bool SELF, int NW, int S, int MINB, int MODE, bool REG = false, bool SD = false, bool LISTED = false>
__global__ void __launch_bounds__(NW * 32, MINB)
cheb_pair_step(const int32_t *__restrict__ dcode, const double *__restrict__ table, const double *__restrict__ dtab,
               const double2 *__restrict__ xa /* T_{n-1} */, const double2 *__restrict__ xb /* T_n */,
               double2 *__restrict__ xc /* T_{n+1} */, double2 *__restrict__ xd /* T_{n+2} */, int n_sites, int n_panels,
               double alpha, double alpha2, double csub, int first, double *__restrict__ partials, unsigned *__restrict__ tickets,
               double *__restrict__ dots_step, const PairWalk wk) {
    // MODE 1: E_{j+1} = alpha2 H~u - (csub E_j + E_{j-1}); csub = 2, or 1 in the first launch, where E_{-1} is never loaded (= 0)
    constexpr int W = NW * S, R = kRecBytes;
    // (lazy hand-over: one guard record behind every T_{n+1} plane -- the first / last warp's outer neighbour read, whose
    // value never matters, must not land in the slot [A] is writing while [B] of the same iteration reads)
    constexpr uint32_t PLANE_N = (W + 2) * R, PLANE_W = (W + kRingGuard) * R;
    extern __shared__ __align__(128) unsigned char pair_smem[];
    const uint32_t sTn = smem_u32(pair_smem);         // T_n planes, local site l2 = y - (y0 - 2)
    const uint32_t sT1 = sTn + kRingN * PLANE_N + R;  // guard record, then T_{n+1} planes (computed here), local site l = y - (y0 - 1)
    const uint32_t sBar = sT1 + kRing * PLANE_W + R;  // guard record (halo rows of [B] read one before / past the ring), then one mbarrier per T_n slot

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // LISTED: the panel of the pieces in hand and the run (= partial-sum slot) they belong to
    int panel = -1, run = -1;

    for (uint32_t o = threadIdx.x * 16u; o < kRingN * PLANE_N + kRing * PLANE_W + 2 * R; o += NW * 32 * 16u)
        sts_rec(sTn + o, make_double2(0.0, 0.0));
    if (threadIdx.x == 0) {
#pragma unroll
        for (int r = 0; r < kRingN; ++r) mbar_init(sBar + 8 * r, 1);
        mbar_init(sBar + 8 * kRingN, NW * 32);  // [A] -> [B] hand-over: every thread arrives, waits later (split phase)
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // the clears are ordered before the bulk copies
    __syncthreads();
    uint32_t cnt = 0;  // T_n planes consumed by the items before this one
#if defined(BDG_PAIR_SPLIT) || defined(BDG_PAIR_LAZY)
    uint32_t xphase = 0;  // parity of the hand-over barrier's current phase (one phase per iteration)
#endif
    const size_t gstep = (size_t)wk.M * 32;
    double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;

    {
    const int l0 = S * warp;  // this warp's sites: l0, l0 + 1
    // own-record addresses of site l0 in slot 0 of each ring (+ lane's 16 bytes)
    uint32_t aN = sTn + (uint32_t)(l0 + 1) * R + (uint32_t)lane * 16u;
    uint32_t a1 = sT1 + (uint32_t)l0 * R + (uint32_t)lane * 16u;
    pin(aN), pin(a1);

    double keep[S][kDirs];
    int jheld = -2;
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int u = 0; u < kDirs; ++u) keep[s][u] = 0.0;
    // lanes 0 .. 5 S - 1 carry the codes of the warp's rows: lane = s * 5 + direction
    const int csite = lane / kDirs, cdir = lane - csite * kDirs;
    const bool code_lane = lane < S * kDirs, self_lane = cdir == 0;
    const int plane_codes = wk.M * kDirs, all_codes = wk.Lx * plane_codes;

    const int item0 = LISTED ? wk.cta_begin[blockIdx.x] : (int)blockIdx.x, item1 = LISTED ? wk.cta_begin[blockIdx.x + 1] : wk.n_items;
    for (int item = item0; item < item1; item += LISTED ? 1 : (int)gridDim.x) {
        int4 piece;  // panel, patch, x0, len
        if (LISTED) {
            piece = wk.pieces[item];  // (run, patch, x0, len)
            if (piece.x != run) {
                if (run >= 0) {
                    flush_dots<NW>(d0, d1, d2, d3, pair_smem, run, wk.panel_runs[panel], wk.panel_runs[panel + 1], panel, n_panels, partials,
                                   tickets, dots_step);
                    d0 = d1 = d2 = d3 = 0.0;
                }
                run = piece.x;
                panel = wk.run_panel[run];
            }
            piece.x = panel;
        } else {
            const int seg = item / wk.n_patches;
            piece.x = (int)blockIdx.y, piece.y = item - seg * wk.n_patches, piece.z = seg * wk.seg_len;
            piece.w = min(wk.Lx, piece.z + wk.seg_len) - piece.z;
        }
        const size_t pbase = (size_t)piece.x * n_sites * 32;
        const double2 *const ta = xa + pbase, *const tb = xb + pbase;
        double2 *const tc = xc + pbase, *const td = xd + pbase;  // MODE 1: td = ta (E_{j+1} over E_{j-1}), tc unused
        const int patch = piece.y;
        // Local site l = 1 of a patch is y0.  On a plane with OPEN ends (wk.open: no block wraps around in-plane) the site
        // below the first patch / above the last one does not exist, so those two patches own their rim site too (l = 0 resp.
        // l = P + 1) instead of recomputing a halo value nobody needs: n P + 2 sites in n patches (100 sites: 7 instead of 8).
        const int y0 = patch * wk.P + wk.open, x0 = piece.z, len = piece.w;
        const int own_lo = wk.open && patch == 0 ? 0 : 1, own_hi = wk.P + (wk.open && patch == wk.n_patches - 1 ? 1 : 0);
        const int n_planes = len + 4;  // T_n planes x0 - 2 .. x0 + len + 1 (mod Lx); plane t -> ring slot (cnt + t) % kRingN
        // In-plane run of a T_n plane: y = y0 - 2 .. ye + 1 (ye = end of the owned sites; at most the W + 2 records of a ring
        // plane), l2 = y - (y0 - 2); the part below 0 / from M on comes from the opposite side of the plane (its own small
        // bulk copy).
        const int ye = min(wk.M, y0 + own_hi), span = min(ye - y0 + 4, W + 2);
        const int n_lo = max(0, 2 - y0), n_hi = max(0, y0 - 2 + span - wk.M);
        const double2 *src_lo = tb + (ptrdiff_t)(y0 - 2 + wk.M) * 32, *src_main = tb + (ptrdiff_t)(y0 - 2 + n_lo) * 32,
                      *src_hi = tb + (ptrdiff_t)(y0 - 2 + span - n_hi - wk.M) * 32;
        // One elected thread stages a plane: expect the bytes on the slot's barrier, then up to three bulk copies.
        auto issue = [&](int t) {
            if (warp == 0 && t < n_planes) {
                int xw = x0 - 2 + t;
                xw += xw < 0 ? wk.Lx : 0;
                xw -= xw >= wk.Lx ? wk.Lx : 0;
                const size_t po = (size_t)xw * gstep;
                const uint32_t r = (cnt + (uint32_t)t) & (kRingN - 1);
                const uint32_t dst = sTn + r * PLANE_N, bar = sBar + 8 * r;
                if (lane == 0) {
                    mbar_expect_tx(bar, (uint32_t)span * R);
                    bulk_g2s(dst + (uint32_t)n_lo * R, src_main + po, (uint32_t)(span - n_lo - n_hi) * R, bar);
                    if (n_lo) bulk_g2s(dst, src_lo + po, (uint32_t)n_lo * R, bar);
                    if (n_hi) bulk_g2s(dst + (uint32_t)(span - n_hi) * R, src_hi + po, (uint32_t)n_hi * R, bar);
                }
            }
        };
        auto wait = [&](int t) {
            if (t < n_planes) {
                const uint32_t c = cnt + (uint32_t)t;
                mbar_wait(sBar + 8 * (c & (kRingN - 1)), (c >> 3) & 1u);
            }
        };

        __syncthreads();  // every warp is done with the previous item's planes
        for (int t = 0; t < kRingN; ++t) issue(t);

        // Site l0 + s of this warp: y = y0 - 1 + l0 + s (mod M).  Owned = inside the patch and the lattice.
        const int ya = y0 - 1 + l0;
        bool owned[S];
#pragma unroll
        for (int s = 0; s < S; ++s) owned[s] = l0 + s >= own_lo && l0 + s <= own_hi && ya + s < wk.M;
        // code index of this lane's (row, direction) in plane 0; rows past the halo (ragged patch) wrap anywhere valid
        int yl = (ya + csite) % wk.M;
        yl += yl < 0 ? wk.M : 0;
        const int coff = yl * kDirs + cdir;
        int cplane = x0 - 1;  // plane of [A] in iteration 0
        cplane += cplane < 0 ? wk.Lx : 0;
        cplane *= plane_codes;
        // codes of the planes of [A] in this iteration, [B] (one plane behind), and of the next two planes of [A]: loaded
        // two planes ahead, so that the on-site fragment fetch of the next plane (SELF) never waits for its code
        int jvA = -1, jvB = -1, jnext = -1;
        if (code_lane) jvA = __ldg(dcode + cplane + coff);
        cplane += plane_codes;
        cplane -= cplane >= all_codes ? all_codes : 0;
        if (code_lane) jnext = __ldg(dcode + cplane + coff);
        double fsA[S], fsB[S], fsN[S];  // SELF: on-site fragments of the planes of [A], [B] and of the next [A]
#pragma unroll
        for (int s = 0; s < S; ++s) fsA[s] = fsB[s] = fsN[s] = 0.0;
        if (SELF) self_fragments<S, SD>(jvA, fsA, table, dtab, lane);
        // Global element of (plane x0 - 1, site ya): T_{n+1} / E_{j-1} / E_{j+1} of the warp's rows; only owned rows of
        // owned planes are dereferenced through it.  The second row sits 32 elements on.
        const ptrdiff_t gbase = ((ptrdiff_t)(x0 - 1) * wk.M + ya) * 32 + lane;
        const double2 *pin_ = ta + gbase;                       // MODE 1: E_{j-1}(plane of [A])
        double2 *pout1 = tc + gbase;                            // MODE 0: T_{n+1}(plane of [A])
        double2 *pout2 = td + gbase - (ptrdiff_t)gstep;         // T_{n+2} / E_{j+1}(plane of [B]) = one plane behind [A]
        // MODE 0: T_{n-1} of ALL rows [A] computes (halo rows and halo planes included), wrapped like the codes;
        // MODE 1: E_{j-1} of the owned rows of the plane [B] works on.  Both one plane ahead (pvn).
        int yw[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            yw[s] = (ya + s) % wk.M;
            yw[s] += yw[s] < 0 ? wk.M : 0;
        }
        int xprev = x0 - 1;  // wrapped plane of [A]
        xprev += xprev < 0 ? wk.Lx : 0;
        double2 pv[S], pvn[S];  // MODE 1: E_{j-1} of the plane of [B] in this / in the next iteration
#ifdef BDG_PAIR_PV3
        double2 pvn2[S];        // ... and in the one after: the load is issued two and a half iterations ahead of its use
#pragma unroll
        for (int s = 0; s < S; ++s) pvn2[s] = make_double2(0.0, 0.0);
#endif
#pragma unroll
        for (int s = 0; s < S; ++s) {
            pv[s] = pvn[s] = make_double2(0.0, 0.0);
            if (MODE == 0) pvn[s] = ld_prev(ta + ((size_t)xprev * wk.M + yw[s]) * 32 + lane);
        }
        wait(0);
        wait(1);
        double2 rt[S], rc[S], o1[S], o2[S];  // REG: own records of T_n planes i, i + 1 and of T_{n+1} planes i - 1, i - 2
#pragma unroll
        for (int s = 0; s < S; ++s) {
            rt[s] = rc[s] = o1[s] = o2[s] = make_double2(0.0, 0.0);
            if (REG) {
                rt[s] = lds_rec(aN + (cnt & (kRingN - 1)) * PLANE_N + (uint32_t)s * R);
                rc[s] = lds_rec(aN + ((cnt + 1) & (kRingN - 1)) * PLANE_N + (uint32_t)s * R);
            }
        }

        for (int i = 0; i <= len + 1; ++i, pin_ += gstep, pout1 += gstep, pout2 += gstep) {
            const bool store = i >= 1 && i <= len;
            if (MODE == 1) {
                // E_{j-1} of the plane of [A]: [B] of the NEXT iteration works on that plane (and overwrites it) -- issued
                // here, an iteration and a half ahead of its use (half an iteration does not cover the HBM latency: measured)
#ifdef BDG_PAIR_PV3
                const bool store_next = i + 1 <= len;  // the plane of [A] in the next iteration is an owned plane
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    pv[s] = pvn[s];
                    pvn[s] = pvn2[s];
                    if (store_next && owned[s] && !first) pvn2[s] = ld_prev_rw(pin_ + gstep + 32 * s);
                }
#else
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    pv[s] = pvn[s];
                    if (store && owned[s] && !first) pvn[s] = ld_prev_rw(pin_ + 32 * s);
                }
#endif
            } else {  // T_{n-1} of the next plane of [A], one iteration ahead
                if (i <= len) {
                    xprev += 1;
                    xprev -= xprev >= wk.Lx ? wk.Lx : 0;
                }
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    pv[s] = pvn[s];
                    if (i <= len) pvn[s] = ld_prev(ta + ((size_t)xprev * wk.M + yw[s]) * 32 + lane);
                }
            }
            int jn2 = -1;
            cplane += plane_codes;
            cplane -= cplane >= all_codes ? all_codes : 0;
            if (code_lane && i < len) jn2 = __ldg(dcode + cplane + coff);
            if (SELF) self_fragments<S, SD>(jnext, fsN, table, dtab, lane);
            double2 tn[S] = {};  // T_n of the warp's rows in the plane of [B] = the records [A] reads as its x-1 neighbours
            double2 out[S] = {};  // T_{n+1} of the warp's rows in the plane of [A]
            // [B]: update, store and dot products of row s from its product (yr, yi) and its own T_{n+1} record
            auto finish_b = [&](int s, double yr, double yi, const double2 &own1) {
                const double2 t = tn[s];
                double2 out;
                if (MODE == 0) {
                    out = make_double2(fma(alpha, yr, -t.x), fma(alpha, yi, -t.y));
                } else {
                    const double2 sub = make_double2(fma(csub, t.x, pv[s].x), fma(csub, t.y, pv[s].y));
                    out = make_double2(fma(alpha2, yr, -sub.x), fma(alpha2, yi, -sub.y));
                }
                if (owned[s]) {
                    pout2[32 * s] = out;
                    if (MODE == 0) {
                        d2 = fma(own1.x, own1.x, fma(own1.y, own1.y, d2));
                    } else {
                        d2 = fma(out.x, t.x, fma(out.y, t.y, d2));
                    }
                    d3 = fma(out.x, own1.x, fma(out.y, own1.y, d3));
                }
            };
            {
                const uint32_t c = cnt + (uint32_t)i;
                const uint32_t nm = aN + (c & (kRingN - 1)) * PLANE_N;
                const uint32_t n0 = aN + ((c + 1) & (kRingN - 1)) * PLANE_N;
                const uint32_t np = aN + ((c + 2) & (kRingN - 1)) * PLANE_N;
                hold_fragments<DIAG, SELF, S, SD>(jvA, jheld, keep, table, dtab, lane, self_lane);
                double2 c_[S + 2], q[S];  // c_[1 + s] = the own record of site s; c_[0], c_[S + 1] = the warp's in-plane neighbours
                if (REG) {
                    c_[0] = lds_rec(n0 - R);
                    c_[S + 1] = lds_rec(n0 + (uint32_t)S * R);
#pragma unroll
                    for (int s = 0; s < S; ++s) c_[1 + s] = rc[s], tn[s] = rt[s];
                } else {
#pragma unroll
                    for (int k = 0; k < S + 2; ++k) c_[k] = lds_rec(n0 + (uint32_t)(k - 1) * R);
#pragma unroll
                    for (int s = 0; s < S; ++s) tn[s] = lds_rec(nm + (uint32_t)s * R);
                }
                const uint32_t t1 = a1 + (uint32_t)(i & (kRing - 1)) * PLANE_W;
                // The newest plane (x+1 neighbours) feeds the LAST term of a row: its barrier is waited for only after the
                // other four terms are under way (+1 %: profiles/r02/21_latewait.log).
                RowSum<DIAG, SD> rs[S];
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    rs[s].begin(c_[1 + s], SELF ? fsA[s] : keep[s][0]);
                    rs[s].add(tn[s], keep[s][1]);
                    rs[s].add(c_[s], keep[s][2]);
                    rs[s].add(c_[2 + s], keep[s][3]);
                }
                wait(i + 2);
#pragma unroll
                for (int s = 0; s < S; ++s) q[s] = lds_rec(np + (uint32_t)s * R);
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    double yr, yi;
                    rs[s].add(q[s], keep[s][4]);
                    rs[s].end();
                    yr = rs[s].yr, yi = rs[s].yi;
                    if (MODE == 0) out[s] = make_double2(fma(alpha, yr, -pv[s].x), fma(alpha, yi, -pv[s].y));
                    else out[s] = make_double2(alpha * yr, alpha * yi);
                    sts_rec(t1 + (uint32_t)s * R, out[s]);
                }
                if (REG) {
#pragma unroll
                    for (int s = 0; s < S; ++s) rt[s] = rc[s], rc[s] = q[s];
                }
                if (store) {
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        if (owned[s]) {
                            if (MODE == 0) pout1[32 * s] = out[s];
                            d0 = fma(c_[1 + s].x, c_[1 + s].x, fma(c_[1 + s].y, c_[1 + s].y, d0));
                            d1 = fma(out[s].x, c_[1 + s].x, fma(out[s].y, c_[1 + s].y, d1));
                        }
                    }
                }
            }
#ifndef BDG_PAIR_SPLIT
#ifdef BDG_PAIR_LAZY
            // What [B] of this iteration reads from OTHER warps is a plane old: the in-plane neighbours of T_{n+1}(x - 1), written
            // in [A] of the previous iteration (the planes x - 2 and x it reads at the warp's own sites only).  So the CTA-wide
            // hand-over it needs is the one of the PREVIOUS iteration: every thread arrives at the end of an iteration and waits
            // for that phase only after [A] of the next one -- half an iteration of slack instead of none.
            if (i >= 1) {
                mbar_wait(sBar + 8 * kRingN, xphase);
                xphase ^= 1u;
                issue(i - 1 + kRingN);  // every warp is through [A] of iteration i - 1: plane i - 1 is dead
            }
#else
            __syncthreads();  // T_{n+1} of this plane complete in the ring
            // ... and [A] is done in every warp: plane i (its x-1 neighbours) is dead, its slot takes plane i + 8 -- seven
            // planes in flight beyond the one in use.
            issue(i + kRingN);
#endif
            if (i >= 2) {
                const uint32_t t0 = a1 + (uint32_t)((i - 1) & (kRing - 1)) * PLANE_W;
                const uint32_t tm = a1 + (uint32_t)((i - 2) & (kRing - 1)) * PLANE_W;
                const uint32_t tp = a1 + (uint32_t)(i & (kRing - 1)) * PLANE_W;
                hold_fragments<DIAG, SELF, S, SD>(jvB, jheld, keep, table, dtab, lane, self_lane);
                double2 c_[S + 2], m[S], q[S];
                if (REG) {
                    c_[0] = lds_rec(t0 - R);
                    c_[S + 1] = lds_rec(t0 + (uint32_t)S * R);
#pragma unroll
                    for (int s = 0; s < S; ++s) c_[1 + s] = o1[s], m[s] = o2[s], q[s] = out[s];
                } else {
#pragma unroll
                    for (int k = 0; k < S + 2; ++k) c_[k] = lds_rec(t0 + (uint32_t)(k - 1) * R);
#pragma unroll
                    for (int s = 0; s < S; ++s) m[s] = lds_rec(tm + (uint32_t)s * R);
#pragma unroll
                    for (int s = 0; s < S; ++s) q[s] = lds_rec(tp + (uint32_t)s * R);
                }
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    double yr, yi;
                    row_product<DIAG, SD>(c_[1 + s], m[s], c_[s], c_[2 + s], q[s], SELF ? fsB[s] : keep[s][0], keep[s], yr, yi);
                    finish_b(s, yr, yi, c_[1 + s]);
                }
            }
#else
            // Split-phase hand-over: arrive, start [B] on the T_{n+1} records this warp wrote itself (own sites of the
            // three planes: the on-site product, the x-1 term, and for the upper site its lower neighbour), and only then
            // wait for the other warps' records (one neighbour below, one above).  The order of the terms of a row is
            // unchanged (self, x-1, y-1, y+1, x+1).
            static_assert(S == 2, "the staged [B] is written for two sites per warp");
            mbar_arrive(sBar + 8 * kRingN);
            if (i >= 2) {
                const uint32_t t0 = a1 + (uint32_t)((i - 1) & (kRing - 1)) * PLANE_W;
                const uint32_t tm = a1 + (uint32_t)((i - 2) & (kRing - 1)) * PLANE_W;
                const uint32_t tp = a1 + (uint32_t)(i & (kRing - 1)) * PLANE_W;
                hold_fragments<DIAG, SELF, S, SD>(jvB, jheld, keep, table, dtab, lane, self_lane);
                double2 own[S], m[S], q[S];
                if (REG) {
#pragma unroll
                    for (int s = 0; s < S; ++s) own[s] = o1[s], m[s] = o2[s], q[s] = out[s];
                } else {
#pragma unroll
                    for (int s = 0; s < S; ++s) own[s] = lds_rec(t0 + (uint32_t)s * R);
#pragma unroll
                    for (int s = 0; s < S; ++s) m[s] = lds_rec(tm + (uint32_t)s * R);
#pragma unroll
                    for (int s = 0; s < S; ++s) q[s] = lds_rec(tp + (uint32_t)s * R);
                }
                RowSum<DIAG, SD> r0, r1;
                r0.begin(own[0], SELF ? fsB[0] : keep[0][0]);
                r1.begin(own[1], SELF ? fsB[1] : keep[1][0]);
                r0.add(m[0], keep[0][1]);
                r1.add(m[1], keep[1][1]);
                r1.add(own[0], keep[1][2]);
                mbar_wait(sBar + 8 * kRingN, xphase);
                issue(i + kRingN);
                const double2 lo = lds_rec(t0 - R), hi = lds_rec(t0 + 2 * R);
                r0.add(lo, keep[0][2]);
                r0.add(own[1], keep[0][3]);
                r1.add(hi, keep[1][3]);
                r0.add(q[0], keep[0][4]);
                r1.add(q[1], keep[1][4]);
                r0.end();
                r1.end();
                finish_b(0, r0.yr, r0.yi, own[0]);
                finish_b(1, r1.yr, r1.yi, own[1]);
            } else {
                mbar_wait(sBar + 8 * kRingN, xphase);
                issue(i + kRingN);
            }
            xphase ^= 1u;
#endif
#ifdef BDG_PAIR_LAZY
            mbar_arrive(sBar + 8 * kRingN);  // this thread's T_{n+1} records of the plane are in the ring, its reads of older planes done
#endif
            if (REG) {
#pragma unroll
                for (int s = 0; s < S; ++s) o2[s] = o1[s], o1[s] = out[s];
            }
            jvB = jvA;
            jvA = jnext;
            jnext = jn2;
#pragma unroll
            for (int s = 0; s < S; ++s) fsB[s] = fsA[s], fsA[s] = fsN[s];
        }
#ifdef BDG_PAIR_LAZY
        mbar_wait(sBar + 8 * kRingN, xphase);  // the last iteration's phase: every arrival has its wait
        xphase ^= 1u;
#endif
        cnt += (uint32_t)n_planes;
    }
    }
    if (LISTED) {
        if (panel >= 0)
            flush_dots<NW>(d0, d1, d2, d3, pair_smem, run, wk.panel_runs[panel], wk.panel_runs[panel + 1], panel, n_panels, partials, tickets,
                           dots_step);
    } else {
        const int p = (int)blockIdx.y, r0 = p * (int)gridDim.x;
        flush_dots<NW>(d0, d1, d2, d3, pair_smem, r0 + (int)blockIdx.x, r0, r0 + (int)gridDim.x, p, n_panels, partials, tickets, dots_step);
    }
}

// *bad = 1 unless every block column of the fixed-width copy is the row itself or one of its four nearest
// neighbours on the (x, in-plane) torus -- the open stencil and the reference's periodic one both qualify.
__global__ void __launch_bounds__(256)
pair_check(int64_t n_slots, int width, int Lx, int M, const int32_t *__restrict__ cidx, int *__restrict__ bad) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_slots) return;
    const int row = (int)(t / width);
    if (torus_direction(row, cidx[t], Lx, M) < 0) *bad = 1;
}

// dcode[row][dir] = dictionary code of the row's block towards (self, x-1, y-1, y+1, x+1), -1 = none.
// Slot 0 of the fixed-width copy is the diagonal block; padding slots point at the row itself.
__global__ void __launch_bounds__(256)
pair_codes(int n_sites, int width, int Lx, int M, const int32_t *__restrict__ cidx, const int32_t *__restrict__ ccode,
           int32_t *__restrict__ dcode, int *__restrict__ wraps /* set to 1 when a block wraps around in-plane */) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_sites) return;
    const int y = row % M;
    int out[kDirs] = {-1, -1, -1, -1, -1};
    for (int u = 0; u < width; ++u) {
        const int c = ccode[(size_t)row * width + u];
        const int d = u == 0 ? 0 : torus_direction(row, cidx[(size_t)row * width + u], Lx, M);
        if (u == 0) out[0] = c;
        else if (d == 1) out[1] = c;
        else if (d == 2) out[2] = c;
        else if (d == 3) out[3] = c;
        else if (d == 4) out[4] = c;
    }
#pragma unroll
    for (int k = 0; k < kDirs; ++k) dcode[(size_t)row * kDirs + k] = out[k];
    if ((y == 0 && out[2] >= 0) || (y == M - 1 && out[3] >= 0)) *wraps = 1;
}

using PairKernel = void (*)(const int32_t *, const double *, const double *, const double2 *, const double2 *, double2 *,
                            double2 *, int, int, double, double, double, int, double *, unsigned *, double *, const PairWalk);

template <int NW, int S, int MINB, bool REG = false, bool LISTED = false>
PairKernel pick_pair_shape(bool diag, bool self, bool t2, bool sd = false) {
    if (sd && !diag) {  // general hopping blocks, real-diagonal on-site blocks
        if (self) return t2 ? cheb_pair_step<false, true, NW, S, MINB, 1, REG, true, LISTED> : cheb_pair_step<false, true, NW, S, MINB, 0, REG, true, LISTED>;
        return t2 ? cheb_pair_step<false, false, NW, S, MINB, 1, REG, true, LISTED> : cheb_pair_step<false, false, NW, S, MINB, 0, REG, true, LISTED>;
    }
    if (self) {
        if (t2) return diag ? cheb_pair_step<true, true, NW, S, MINB, 1, REG, false, LISTED> : cheb_pair_step<false, true, NW, S, MINB, 1, REG, false, LISTED>;
        return diag ? cheb_pair_step<true, true, NW, S, MINB, 0, REG, false, LISTED> : cheb_pair_step<false, true, NW, S, MINB, 0, REG, false, LISTED>;
    }
    if (t2) return diag ? cheb_pair_step<true, false, NW, S, MINB, 1, REG, false, LISTED> : cheb_pair_step<false, false, NW, S, MINB, 1, REG, false, LISTED>;
    return diag ? cheb_pair_step<true, false, NW, S, MINB, 0, REG, false, LISTED> : cheb_pair_step<false, false, NW, S, MINB, 0, REG, false, LISTED>;
}

int env_int(const char *name, int fallback) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : fallback;
}

struct PairShape {
    int warps, sites;  // NW, S
    PairKernel kernel;
    PairKernel listed;  // the same kernel working through the lists of a balanced plan (8-warp shape only), else null
    size_t smem;
};

PairShape pair_shape(bool diag, bool self, bool t2, bool sd) {
    PairShape s;
    // 8 warps x 2 sites: two CTAs per SM, one computes while the other waits at its barrier.  The shape
    // sweep (profiles/r01/s4_pair_shape_sweep.log: 16 x 2 one CTA per SM -2 %, 8 x 3 -11 %, 12 x 1 with 24
    // warps per SM -14 %, 6 x 2 with three CTAs per SM -15 %) left this one ahead; 16 x 2 is kept for the tests.
    const int warps = env_int("BDG_PAIR_WARPS", 8);
    s.listed = nullptr;
    if (warps <= 8)
        s.warps = 8, s.sites = 2, s.kernel = pick_pair_shape<8, 2, 2>(diag, self, t2, sd), s.listed = pick_pair_shape<8, 2, 2, false, true>(diag, self, t2, sd);
    else if (warps <= 12)  // one CTA per SM, 168 registers: the warp's own records stay in registers (REG)
        s.warps = 12, s.sites = 2, s.kernel = pick_pair_shape<12, 2, 1, true>(diag, self, t2, sd);
    else
        s.warps = 16, s.sites = 2, s.kernel = pick_pair_shape<16, 2, 1>(diag, self, t2, sd);
    const int W = s.warps * s.sites;
    s.smem = ((size_t)kRingN * (W + 2) + (size_t)kRing * (W + kRingGuard) + 2) * kRecBytes + 8 * (kRingN + 1);
    return s;
}

// On-site fragments per row straight from the table (SELF) when the dictionary is large: then the on-site blocks
// differ from site to site (disorder, a self-consistent gap, a phase winding) and holding them in registers from
// plane to plane would reload them on every row through the slow path.
// Real-diagonal on-site blocks next to general hopping blocks: two multiplications instead of two MMAs (BDG_ELL_SD=0: off)
bool pair_sd(const EllDev &e) { return e.self_diag_usable && !e.diag_usable && env_int("BDG_ELL_SD", 1) != 0; }

bool pair_self(const EllDev &e) {
    const int forced = env_int("BDG_PAIR_SELF", -1);
    return forced >= 0 ? forced != 0 : e.n_unique > 64;
}

}  // namespace

bool pair_streams_onsite(const bdg_system *sys) { return pair_self(sys->ell); }

// Does the current fixed-width copy qualify?  (dictionary built, one-dimensional x-planes of >= 3 sites, >= 3
// planes, nearest-neighbour stencil -- open or periodic --, rows of <= 5 blocks)
int pair_probe(bdg_system *sys) {
    EllDev &e = sys->ell;
    e.pair_usable = false;
    e.pair_M = 0;
    if (!e.usable || !e.dict_usable || e.width > 5 || e.width < 3) return BDG_OK;
    const int Lx = sys->cubic[0], Ly = sys->cubic[1], Lz = sys->cubic[2];
    if ((int64_t)Lx * Ly * Lz != e.n_sites) return BDG_OK;
    const int M = Lz == 1 ? Ly : (Ly == 1 ? Lz : 0);
    if (M < 3 || Lx < 3) return BDG_OK;
    BDG_TRY(ensure_scratch(sys, 2, 64));
    int *bad = sys->scratch_i32[2].as<int>();
    BDG_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), sys->stream));
    const int64_t n_slots = e.n_sites * e.width;
    pair_check<<<(unsigned)ceil_div(n_slots, 256), 256, 0, sys->stream>>>(n_slots, e.width, Lx, M, e.idx.as<int32_t>(), bad);
    BDG_CUDA(cudaGetLastError());
    int host = 1;
    BDG_CUDA(cudaMemcpyAsync(&host, bad, sizeof(int), cudaMemcpyDeviceToHost, sys->stream));
    BDG_CUDA(cudaStreamSynchronize(sys->stream));
    e.pair_usable = host == 0;
    e.pair_M = M;
    if (e.pair_usable) {
        BDG_TRY(dev_alloc(sys, e.dcode, (size_t)e.n_sites * kDirs * sizeof(int32_t)));
        pair_codes<<<(unsigned)ceil_div(e.n_sites, 256), 256, 0, sys->stream>>>((int)e.n_sites, e.width, Lx, M, e.idx.as<int32_t>(),
                                                                              e.code.as<int32_t>(), e.dcode.as<int32_t>(), bad);
        BDG_CUDA(cudaGetLastError());
        // (`bad` is 0 here.)  Open plane ends let the rim patches own their rim site; the flag stays valid under incremental
        // updates, which cannot add a block (a changed zero pattern rebuilds everything).
        BDG_CUDA(cudaMemcpyAsync(&host, bad, sizeof(int), cudaMemcpyDeviceToHost, sys->stream));
        BDG_CUDA(cudaStreamSynchronize(sys->stream));
        e.pair_open = host == 0;
    }
    return BDG_OK;
}

// Patch size, segment length and grid for the current recursion.
int pair_configure(bdg_system *sys) {
    ChebState &st = sys->cheb;
    const EllDev &e = sys->ell;
    const PairShape shape = pair_shape(st.kernel == BDG_KERNEL_DICT_DIAG, pair_self(e), st.t2, pair_sd(e));
    BDG_CUDA(cudaFuncSetAttribute(shape.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shape.smem));
    if (shape.listed) BDG_CUDA(cudaFuncSetAttribute(shape.listed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shape.smem));
    int per_sm = 1;
    BDG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, shape.kernel, shape.warps * 32, shape.smem));
    per_sm = std::max(per_sm, 1);
    const int64_t slots = std::max<int64_t>(1, (int64_t)sys->sm_count * per_sm / st.n_panels);
    PairWalk &w = st.pair_walk;
    w.Lx = sys->cubic[0];
    w.M = e.pair_M;
    const int p_max = shape.warps * shape.sites - 2;
    w.open = e.pair_open && env_int("BDG_PAIR_OPEN", 1) != 0 ? 1 : 0;
    const int to_cover = std::max(1, w.M - 2 * w.open);  // n patches own n P (+ 2: the rim sites of a plane with open ends) sites
    const int n_patches_min = (int)ceil_div(to_cover, p_max);
    w.P = (int)ceil_div(to_cover, n_patches_min);  // balanced patches
    w.P = std::max(1, std::min(env_int("BDG_PAIR_P", w.P), p_max));
    w.n_patches = (int)ceil_div(to_cover, w.P);
    // Segment length: every item recomputes one plane of T_{n+1} on either side of its segment and
    // starts with a cold pipeline (about two more plane times); many items balance the CTAs.
    double best = -1.0;
    const int forced = env_int("BDG_PAIR_SEG", 0);
    for (int n_seg = 1; n_seg <= std::max(1, w.Lx / 4); ++n_seg) {
        const int len = forced > 0 ? std::min(forced, w.Lx) : (int)ceil_div(w.Lx, n_seg);
        const int64_t segs = ceil_div(w.Lx, len);
        const int64_t items = (int64_t)w.n_patches * segs;
        const double balance = (double)items / (double)(ceil_div(items, slots) * slots);
        const double score = balance * len / (len + 4.0);
        if (score > best) {
            best = score;
            w.seg_len = len;
            w.n_segs = (int)segs;
            w.n_items = (int)items;
        }
        if (forced > 0) break;
    }
    // ---- classic or balanced plan ------------------------------------------------------------------------------------
    // Classic: every panel has its own `gx` CTAs (grid = (gx, panels)), CTA (bx, panel) takes the items bx, bx + gx, ...
    // (segment-major: the grid sweeps the lattice as one wavefront, so the halo sites of a patch are in L2 when its neighbour
    // reads them); the kernel computes its items from the item index.
    // Balanced: the (panel, patch, x) space cut into one contiguous chunk per CTA slot, handed to the kernel as lists -- no
    // wave quantisation, no per-panel slot quantisation, fewer segment halos.  Taken when its longest CTA is clearly shorter
    // (small lattices with many panels: C2 / C3: +3 .. +17 %); on large lattices the classic plan is near-perfect already.
    constexpr int kPieceCost = 6;  // iterations a piece costs beyond its planes: two halo planes either side, cold pipeline
    const int gx = (int)std::min<int64_t>(slots, w.n_items);
    const int64_t all_slots = (int64_t)sys->sm_count * per_sm;
    const double classic_cost = (double)ceil_div((int64_t)gx * st.n_panels, all_slots) * (double)ceil_div(w.n_items, gx) * (w.seg_len + kPieceCost);
    w.pieces = nullptr, w.cta_begin = w.run_panel = w.panel_runs = nullptr;
    w.n_ctas = gx, w.n_runs = gx * st.n_panels;
    st.pair_grid_x = gx;
    if (!shape.listed) return BDG_OK;
    const WorkPlan plan = best_balanced_plan(st.n_panels, w.n_patches, w.Lx, all_slots, kPieceCost, 1.08);
    const int force = env_int("BDG_PAIR_BALANCE", -1);
    if (!(force >= 0 ? force != 0 : plan.longest < 0.93 * classic_cost)) return BDG_OK;
    WorkLists lists;
    BDG_TRY(upload_work_lists(sys, st.work_items, plan, st.n_panels, lists));
    w.pieces = lists.pieces, w.cta_begin = lists.cta_begin, w.run_panel = lists.run_panel, w.panel_runs = lists.panel_runs;
    w.n_ctas = lists.n_ctas, w.n_runs = lists.n_runs;
    st.pair_grid_x = (int)ceil_div(lists.n_runs, st.n_panels);  // (sizes the partial-sum buffer: n_panels x pair_grid_x runs)
    return BDG_OK;
}

// Pair mode: T_{n+1} -> x_next1, T_{n+2} -> x_next2 and the dot products of both steps (dots_step: 4 rows).
int pair_launch(bdg_system *sys, const void *x_prev, const void *x_cur, void *x_next1, void *x_next2, double *dots_step) {
    ChebState &st = sys->cheb;
    const EllDev &e = sys->ell;
    const PairShape shape = pair_shape(st.kernel == BDG_KERNEL_DICT_DIAG, pair_self(e), false, pair_sd(e));
    const bool listed = st.pair_walk.pieces != nullptr;
    dim3 grid((unsigned)st.pair_walk.n_ctas, listed ? 1u : (unsigned)st.n_panels);
    (listed ? shape.listed : shape.kernel)<<<grid, shape.warps * 32, shape.smem, sys->stream>>>(
        e.dcode.as<int32_t>(), e.table.as<double>(), e.dtab.as<double>(), static_cast<const double2 *>(x_prev),
        static_cast<const double2 *>(x_cur), static_cast<double2 *>(x_next1), static_cast<double2 *>(x_next2),
        (int)e.n_sites, st.n_panels, 2.0 / st.scale, 0.0, 0.0, 0, st.partials.as<double>(), st.tickets.as<unsigned>(), dots_step,
        st.pair_walk);
    BDG_CUDA(cudaGetLastError());
    return BDG_OK;
}

// T2 mode: E_{j+1} = 2 T_2(H~) E_j - E_{j-1} written over E_{j-1} (x_io); first: E_1 = T_2(H~) E_0.
int t2_launch(bdg_system *sys, bool first, const void *x_cur, void *x_io, double *dots_step) {
    ChebState &st = sys->cheb;
    if (st.cube) return cube_launch(sys, first, x_cur, x_io, dots_step);
    const EllDev &e = sys->ell;
    const PairShape shape = pair_shape(st.kernel == BDG_KERNEL_DICT_DIAG, pair_self(e), true, pair_sd(e));
    const bool listed = st.pair_walk.pieces != nullptr;
    dim3 grid((unsigned)st.pair_walk.n_ctas, listed ? 1u : (unsigned)st.n_panels);
    (listed ? shape.listed : shape.kernel)<<<grid, shape.warps * 32, shape.smem, sys->stream>>>(
        e.dcode.as<int32_t>(), e.table.as<double>(), e.dtab.as<double>(), static_cast<const double2 *>(x_io),
        static_cast<const double2 *>(x_cur), nullptr, static_cast<double2 *>(x_io), (int)e.n_sites, st.n_panels,
        1.0 / st.scale, (first ? 2.0 : 4.0) / st.scale, first ? 1.0 : 2.0, first ? 1 : 0, st.partials.as<double>(), st.tickets.as<unsigned>(),
        dots_step, st.pair_walk);
    BDG_CUDA(cudaGetLastError());
    return BDG_OK;
}
