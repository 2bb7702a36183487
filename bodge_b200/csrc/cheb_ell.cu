// Chebyshev step on the kernel-native matrix format ("ELL"): fixed-width block rows, diagonal
// block first, blocks stored in MMA B-fragment order.  This is the default step kernel for
// lattice Hamiltonians (every row has <= 8 blocks and almost all rows have the same count);
// ragged / long-row matrices use the generic BSR kernel in cheb.cu.
//
// Formulation.  For block B = Br + i Bi (4x4) and record X = Xr + i Xi (4 components x PW columns)
//     Y^T = X^T B^T :   acc1 = Xr^T * Bop,  acc2 = Xi^T * Bop,   Bop[b][2a]   = Br[a][b]
//                                                               Bop[b][2a+1] = Bi[a][b]
// so that acc1 = (RR[a], IR[a]) and acc2 = (RI[a], II[a]) land in the SAME lane and
//     Re y[a] = acc1.0 - acc2.1,   Im y[a] = acc1.1 + acc2.0
// need no shuffle.  With mma.m8n8k4 (A 8x4 row, B 4x8 col, C 8x8):
//   A fragment: lane l = X[component l%4][column l/4]   = element l of the site record
//   B fragment: lane l = Bop[l%4][l/4]                  = double l of the stored block
//   C fragment: lane l = (column l/4, n = 2(l%4), 2(l%4)+1) -> y[component l%4][column l/4]
// i.e. a lane's output element is the record element it loaded: T_{n+1}[row] is stored, and
// T_{n-1}[row] loaded, with the same fully coalesced 128-bit access, and T_n[row] (needed by the
// dot products) IS the record already fetched for the diagonal block in slot 0.
//
// Per row and panel the warp issues CH 64-bit block loads (one pass over the matrix serves up
// to NP panels: the B fragments stay in registers), CH + 1 128-bit record loads and one 128-bit
// store -- about half the L1 wavefronts of the interleaved-complex formulation in cheb.cu, which
// ncu showed to be the co-limiter next to HBM (l1tex data-pipe wavefronts 77 % at 1.9 GHz).
#include <algorithm>
#include <cstdlib>

#include "bdg_internal.h"
#include "cheb_device.cuh"

namespace {

__device__ __forceinline__ uint64_t l2_policy(bool evict_first) {
    uint64_t p;
    if (evict_first)
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
    else
        asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}

// Matrix blocks: read once per pass, never from L1 again.
__device__ __forceinline__ double ld_block(const double *p, uint64_t policy) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;\n" : "=d"(v) : "l"(p), "l"(policy));
    return v;
}

template <int PW, int CH, int NP, int PB>
__global__ void __launch_bounds__(kThreads, 4)
cheb_step_ell(const int32_t *__restrict__ cidx, const double *__restrict__ cdata, const double2 *__restrict__ x_cur,
              double2 *__restrict__ x_io, int n_sites, int n_panels, double alpha, double beta, int first,
              int stream_matrix, double *__restrict__ partials, unsigned *__restrict__ tickets,
              double *__restrict__ dots_step) {
    constexpr int REC = PW * 4;           // complex elements per site record
    static_assert(NP % PB == 0, "PB = panels whose loads are in flight together");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int panel0 = blockIdx.y * NP;
    const size_t plane = (size_t)n_sites * REC;
    const bool x_lane = lane < REC;       // lanes past the record (PW < 8) compute on element 0 and discard
    const int x_elem = x_lane ? lane : 0;
    const uint64_t policy = l2_policy(stream_matrix != 0);
    // Wavefront traversal (see cheb.cu): in pass t the grid works on rows (t*gridDim.x + b)*kWarps + w.
    const int stride = gridDim.x * kWarps;

    int row = blockIdx.x * kWarps + warp;
    int jv = 0;
    if (row < n_sites && lane < CH) jv = __ldg(cidx + (size_t)row * CH + lane);

    double d0[NP], d1[NP];
#pragma unroll
    for (int pp = 0; pp < NP; ++pp) d0[pp] = d1[pp] = 0.0;

    for (; row < n_sites; row += stride) {
        int jn[CH];
#pragma unroll
        for (int u = 0; u < CH; ++u) jn[u] = __shfl_sync(kFull, jv, u);
        double bop[CH];
        const double *blk = cdata + (size_t)row * CH * 32 + lane;
#pragma unroll
        for (int u = 0; u < CH; ++u) bop[u] = ld_block(blk + u * 32, policy);
        const int nrow = row + stride;
        int jnext = 0;
        if (nrow < n_sites && lane < CH) jnext = __ldg(cidx + (size_t)nrow * CH + lane);
        const size_t off = (size_t)row * REC + x_elem;

#pragma unroll
        for (int pb = 0; pb < NP; pb += PB) {
            double2 xv[PB][CH], pv[PB];
            bool on[PB];
#pragma unroll
            for (int pp = 0; pp < PB; ++pp) {
                const int panel = panel0 + pb + pp;
                on[pp] = NP == 1 || panel < n_panels;          // uniform; ragged last group only
                const size_t base = (size_t)(on[pp] ? panel : n_panels - 1) * plane;
#pragma unroll
                for (int u = 0; u < CH; ++u) xv[pp][u] = ld_reuse(x_cur + base + (size_t)jn[u] * REC + x_elem);
                pv[pp] = make_double2(0.0, 0.0);
                if (!first) pv[pp] = ld_plain(x_io + base + off);
            }
#pragma unroll
            for (int pp = 0; pp < PB; ++pp) {
                double a10 = 0.0, a11 = 0.0, a20 = 0.0, a21 = 0.0;
#pragma unroll
                for (int u = 0; u < CH; ++u) {
                    dmma_8x8x4(a10, a11, xv[pp][u].x, bop[u]);
                    dmma_8x8x4(a20, a21, xv[pp][u].y, bop[u]);
                }
                const double yr = a10 - a21, yi = a11 + a20;
                if (on[pp] && x_lane) {
                    const double2 tn = xv[pp][0];  // slot 0 is the row's own record
                    const double2 out = make_double2(alpha * yr - beta * pv[pp].x, alpha * yi - beta * pv[pp].y);
                    x_io[(size_t)(panel0 + pb + pp) * plane + off] = out;
                    d0[pb + pp] += tn.x * tn.x + tn.y * tn.y;
                    d1[pb + pp] += out.x * tn.x + out.y * tn.y;
                }
            }
        }
        jv = jnext;
    }
    // a column's four components sit in one quad of lanes
#pragma unroll
    for (int pp = 0; pp < NP; ++pp) {
        d0[pp] += __shfl_xor_sync(kFull, d0[pp], 1);
        d1[pp] += __shfl_xor_sync(kFull, d1[pp], 1);
        d0[pp] += __shfl_xor_sync(kFull, d0[pp], 2);
        d1[pp] += __shfl_xor_sync(kFull, d1[pp], 2);
    }
#pragma unroll
    for (int pp = 0; pp < NP; ++pp)
        if (NP == 1 || panel0 + pp < n_panels)
            finish_dots<PW>(d0[pp], d1[pp], lane >> 2, x_lane && (lane & 3) == 0, panel0 + pp, n_panels, partials,
                            tickets, dots_step);
}

// ---- building the format ----------------------------------------------------------------------
// need[row] = blocks of the row, +1 if the diagonal block is absent (slot 0 is reserved for it)
__global__ void __launch_bounds__(256)
ell_row_need(int n_sites, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
             int *__restrict__ max_need) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    int need = 0;
    if (row < n_sites) {
        const int p0 = indptr[row], p1 = indptr[row + 1];
        bool self = false;
        for (int p = p0; p < p1; ++p) self |= indices[p] == row;
        need = p1 - p0 + (self ? 0 : 1);
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) need = max(need, __shfl_xor_sync(0xffffffffu, need, d));
    if ((threadIdx.x & 31) == 0 && need > 0) atomicMax(max_need, need);
}

// One warp per (row, slot); lane l writes double l of the slot in B-fragment order.
__global__ void __launch_bounds__(256)
ell_fill(int n_sites, int width, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
         const double *__restrict__ data, int32_t *__restrict__ cidx, double *__restrict__ cdata) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= (int64_t)n_sites * width) return;
    const int row = (int)(w / width), slot = (int)(w % width);
    const int p0 = indptr[row], cnt = indptr[row + 1] - p0;
    int self = -1;  // position of the diagonal block inside the row
    for (int t = 0; t < cnt; ++t)
        if (indices[p0 + t] == row) self = t;
    int src;  // position inside the row feeding this slot, or -1
    if (slot == 0) {
        src = self;
    } else {
        src = slot - 1;
        if (self >= 0 && src >= self) src += 1;
        if (src >= cnt) src = -1;
    }
    double v = 0.0;
    if (src >= 0) {
        const int a = lane >> 3, part = (lane >> 2) & 1, b = lane & 3;
        v = data[((size_t)(p0 + src) * 16 + a * 4 + b) * 2 + part];
    }
    cdata[w * 32 + lane] = v;
    if (lane == 0) cidx[w] = src >= 0 ? indices[p0 + src] : row;
}

using EllKernel = void (*)(const int32_t *, const double *, const double2 *, double2 *, int, int, double, double, int,
                           int, double *, unsigned *, double *);

template <int PW, int NP, int PB> EllKernel pick_ch(int width) {
    switch (width) {
        case 3: return cheb_step_ell<PW, 3, NP, PB>;
        case 4: return cheb_step_ell<PW, 4, NP, PB>;
        case 5: return cheb_step_ell<PW, 5, NP, PB>;
        case 6: return cheb_step_ell<PW, 6, NP, PB>;
        case 7: return cheb_step_ell<PW, 7, NP, PB>;
        default: return cheb_step_ell<PW, 8, NP, PB>;
    }
}

EllKernel pick_ell(int pw, int np, int pb, int width) {
    switch (pw) {
        case 1: return pick_ch<1, 1, 1>(width);
        case 2: return pick_ch<2, 1, 1>(width);
        case 4: return pick_ch<4, 1, 1>(width);
        default:
            if (np >= 8) return pick_ch<8, 8, 1>(width);
            if (np >= 4) return pb >= 2 ? pick_ch<8, 4, 2>(width) : pick_ch<8, 4, 1>(width);
            if (np >= 2) return pb >= 2 ? pick_ch<8, 2, 2>(width) : pick_ch<8, 2, 1>(width);
            return pick_ch<8, 1, 1>(width);
    }
}

int env_int(const char *name, int fallback) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : fallback;
}

}  // namespace

void ell_release(bdg_system *sys) {
    dev_free(sys, sys->ell.idx);
    dev_free(sys, sys->ell.data);
    sys->ell = EllDev();
}

int ell_build(bdg_system *sys) {
    EllDev &e = sys->ell;
    if (e.valid) return BDG_OK;
    BDG_TRY(build_packed(sys));
    const BsrDev &m = sys->packed;
    const int n = (int)m.n_sites;
    e.usable = false;
    e.n_sites = n;
    if (n > 0) {
        BDG_TRY(ensure_scratch(sys, 2, 64));
        int *max_need = sys->scratch_i32[2].as<int>();
        BDG_CUDA(cudaMemsetAsync(max_need, 0, sizeof(int), sys->stream));
        ell_row_need<<<(unsigned)ceil_div(n, 256), 256, 0, sys->stream>>>(n, m.indptr.as<int32_t>(),
                                                                         m.indices.as<int32_t>(), max_need);
        BDG_CUDA(cudaGetLastError());
        int need = 0;
        BDG_CUDA(cudaMemcpyAsync(&need, max_need, sizeof(int), cudaMemcpyDeviceToHost, sys->stream));
        BDG_CUDA(cudaStreamSynchronize(sys->stream));
        const int width = std::max(need, 3);
        // Padding costs matrix traffic: accept it when small, or when the matrix is too small to matter.
        const bool cheap = (int64_t)n * width * 4 <= m.n_blocks * 5 || (int64_t)n * width <= (1 << 16);
        if (width <= 8 && cheap) {
            e.width = width;
            BDG_TRY(dev_alloc(sys, e.idx, (size_t)n * width * sizeof(int32_t)));
            BDG_TRY(dev_alloc(sys, e.data, (size_t)n * width * 32 * sizeof(double)));
            const int64_t threads = (int64_t)n * width * 32;
            ell_fill<<<(unsigned)ceil_div(threads, 256), 256, 0, sys->stream>>>(
                n, width, m.indptr.as<int32_t>(), m.indices.as<int32_t>(), m.data.as<double>(), e.idx.as<int32_t>(),
                e.data.as<double>());
            BDG_CUDA(cudaGetLastError());
            e.usable = true;
        }
    }
    e.valid = true;
    return BDG_OK;
}

// Panels per group and grid size for the current recursion (ChebState::panel_width / n_panels).
int ell_configure(bdg_system *sys) {
    ChebState &st = sys->cheb;
    const EllDev &e = sys->ell;
    // Measured (profiles/r01/sweep_ell_v1.log): these kernels are latency-bound once the vectors
    // dominate, so occupancy (24 warps/SM at one panel per warp-row) beats reusing the block
    // registers for 4 or 8 panels (16 warps/SM) -- concurrent panel groups sweep the lattice in
    // step and L2 de-duplicates their matrix reads.  Two panels per pass pay off when the matrix
    // is L2-resident or the rows are long (3-D lattices: more block bytes per record byte).
    const size_t matrix_bytes = (size_t)e.n_sites * e.width * 260;
    const bool pair = st.panel_width == 8 && st.n_panels >= 2 && (e.width >= 6 || matrix_bytes < ((size_t)64 << 20));
    st.panels_per_group = pair ? 2 : 1;
    st.panel_batch = st.panels_per_group;
    // tuning overrides (development): BDG_ELL_NP in {1,2,4,8}, BDG_ELL_PB in {1,2}
    if (st.panel_width == 8) {
        st.panels_per_group = std::min(env_int("BDG_ELL_NP", st.panels_per_group), std::max(st.n_panels, 1));
        st.panels_per_group = st.panels_per_group >= 8 ? 8 : st.panels_per_group >= 4 ? 4 : st.panels_per_group >= 2 ? 2 : 1;
        st.panel_batch = st.panels_per_group >= 8 ? 1 : std::min(env_int("BDG_ELL_PB", st.panel_batch), st.panels_per_group);
    }
    st.n_groups = (int)ceil_div(st.n_panels, st.panels_per_group);
    int per_sm = 1;
    BDG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &per_sm, pick_ell(st.panel_width, st.panels_per_group, st.panel_batch, e.width), kThreads, 0));
    per_sm = std::max(per_sm, 1);
    int64_t gx = std::max<int64_t>(1, (int64_t)sys->sm_count * per_sm / st.n_groups);
    gx = std::min<int64_t>(gx, ceil_div(e.n_sites, kWarps));
    st.grid_x = (int)gx;
    return BDG_OK;
}

int ell_launch_step(bdg_system *sys, bool first, const void *x_cur, void *x_io, double *dots_step) {
    ChebState &st = sys->cheb;
    const EllDev &e = sys->ell;
    EllKernel k = pick_ell(st.panel_width, st.panels_per_group, st.panel_batch, e.width);
    // Stream the matrix through L2 (evict-first) only when nothing will read it again soon: one
    // group per pass and a matrix that cannot stay resident in the 126 MB L2 anyway.
    const size_t matrix_bytes = (size_t)e.n_sites * e.width * 260;
    const int stream_matrix = st.n_groups == 1 && matrix_bytes > (size_t)64 << 20;
    dim3 grid((unsigned)st.grid_x, (unsigned)st.n_groups);
    k<<<grid, kThreads, 0, sys->stream>>>(e.idx.as<int32_t>(), e.data.as<double>(), static_cast<const double2 *>(x_cur),
                                          static_cast<double2 *>(x_io), (int)e.n_sites, st.n_panels,
                                          (first ? 1.0 : 2.0) / st.scale, first ? 0.0 : 1.0, first ? 1 : 0,
                                          stream_matrix, st.partials.as<double>(), st.tickets.as<unsigned>(), dots_step);
    BDG_CUDA(cudaGetLastError());
    return BDG_OK;
}
