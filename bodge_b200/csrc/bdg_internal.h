// Internal declarations shared by the libbdg translation units (not part of the ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "bdg.h"

// ---- error plumbing ---------------------------------------------------------------------
void bdg_set_error(const char *fmt, ...);

#define BDG_CUDA(expr)                                                                      \
    do {                                                                                    \
        cudaError_t err__ = (expr);                                                         \
        if (err__ != cudaSuccess) {                                                         \
            bdg_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,             \
                          cudaGetErrorString(err__));                                       \
            return BDG_E_CUDA;                                                              \
        }                                                                                   \
    } while (0)

#define BDG_TRY(expr)                                                                       \
    do {                                                                                    \
        int rc__ = (expr);                                                                  \
        if (rc__ != BDG_OK) return rc__;                                                    \
    } while (0)

#define BDG_REQUIRE(cond, ...)                                                              \
    do {                                                                                    \
        if (!(cond)) {                                                                      \
            bdg_set_error(__VA_ARGS__);                                                     \
            return BDG_E_INVALID;                                                           \
        }                                                                                   \
    } while (0)

// ---- device buffers -----------------------------------------------------------------------
// Thin owner of one cudaMalloc'd array; tracks bytes per handle.
struct DevBuf {
    void *ptr = nullptr;
    size_t bytes = 0;
    template <class T> T *as() const { return static_cast<T *>(ptr); }
};

struct bdg_system;
int dev_alloc(bdg_system *sys, DevBuf &buf, size_t bytes);
void dev_free(bdg_system *sys, DevBuf &buf);

// ---- BSR structure on the device ---------------------------------------------------------
// data layout == scipy bsr_matrix.data: [nb][4][4] complex128 (re, im interleaved), 256 B/block.
struct BsrDev {
    int64_t n_sites = 0;
    int64_t n_blocks = 0;
    DevBuf indptr;   // int32 [n_sites + 1]
    DevBuf indices;  // int32 [n_blocks]   ascending within a row
    DevBuf brow;     // int32 [n_blocks]   block row of every block (skeleton only)
    DevBuf data;     // double2 [n_blocks * 16]
};

// Row traversal of the fixed-width step kernels (cheb_ell.cu).  The lattice is cut into work
// items = (patch of Pz x Py sites in the z-y plane, one site per warp of the CTA) x (segment of
// `seg_len` consecutive x); a CTA marches its patch along x, so the +-x neighbour records of a row
// are the records the same warp touched one and two steps earlier, and the +-y / +-z neighbours
// belong to the warps next to it: they are served by the SM's L1, not by L2.  Items are handed
// out segment-major (item = it * gridDim.x + blockIdx.x), so the grid still sweeps the lattice as
// one wavefront and every vector byte leaves HBM once.  Matrices without cubic geometry use the
// degenerate walk Lz = n_sites, Pz = warps per CTA, Lx = 1 (plain wavefront over consecutive rows).
struct RowWalk {
    int Lz = 1, Ly = 1, Lx = 1;  // extents (site = z + y*Lz + x*Ly*Lz)
    int Pz = 1, Py = 1;          // patch extents; Pz * Py = warps per CTA
    int nPz = 1, n_patches = 1;  // patches along z, patches in the plane
    int seg_len = 1, n_items = 1;
    int prefetch = 0;            // rows ahead (a multiple of Ly*Lz) whose records are prefetched into L1; 0 = off
};

// Item grid of the two-steps-per-pass kernel (cheb_pair.cu): x-planes of M sites, cut into patches of
// P owned sites; an item = (patch, segment of seg_len consecutive x), handed out segment-major.
struct PairWalk {
    int Lx = 1, M = 1;
    int open = 0;  // the plane has open ends (no in-plane wrap-around block): the rim patches own their rim site
    int P = 1, n_patches = 1;
    int seg_len = 1, n_segs = 1, n_items = 1;
    // Work lists (device, ChebState::work_items; work_lists.h): CTA c works through pieces[cta_begin[c] .. cta_begin[c + 1]), a
    // piece = (run, patch, x0, len).  A maximal sequence of pieces of one panel inside a CTA is a "run": its dot products go to
    // partials[run]; run_panel[run] = its panel, and the runs of panel p are panel_runs[p] .. panel_runs[p + 1].
    const int4 *pieces = nullptr;
    const int *cta_begin = nullptr, *run_panel = nullptr, *panel_runs = nullptr;
    int n_ctas = 0, n_runs = 0;
};

// Item grid of the two-applications-per-pass kernel for three-dimensional lattices (cheb_cube.cu): the (y, z) plane
// is cut into nPy x nPz patches; an item = (patch, segment of seg_len consecutive x), handed out segment-major.
struct CubeWalk {
    int Lx = 1, Ly = 1, Lz = 1;
    int nPy = 1, nPz = 1, n_patches = 1;
    int seg_len = 1, n_segs = 1, n_items = 1;
    // work lists of a balanced plan (see PairWalk), null for the classic plan
    const int4 *pieces = nullptr;
    const int *cta_begin = nullptr, *run_panel = nullptr, *panel_runs = nullptr;
    int n_ctas = 0, n_runs = 0;
};

// ---- Chebyshev state ----------------------------------------------------------------------
struct ChebState {
    bool active = false;
    int kernel = BDG_KERNEL_DMMA;
    int32_t n_cols = 0;       // user columns on this GPU
    int32_t panel_width = 8;  // PW: columns per panel (1, 2, 4 or 8)
    int32_t n_panels = 0;
    double scale = 1.0;
    int32_t steps_done = 0;   // recursion steps after T_1 (T_{steps_done+1} is current)
    int32_t dot_capacity = 0; // steps for which dot storage exists
    // Vectors: [panel][site][col_in_panel][alpha] complex128; cur = T_n, prev = T_{n-1}.
    // The pair kernel writes T_{n+1}, T_{n+2} into the two buffers that hold neither (vec[2], vec[3]
    // exist only then); the single-step kernels overwrite prev in place.
    DevBuf vec[4];
    int cur = 0, prev = 1;
    bool pair = false;         // two steps per launch whenever two or more remain (cheb_pair.cu)
    bool t2 = false;           // even-vector recursion E_{j+1} = 2 T_2(H~) E_j - E_{j-1}: vec[cur] = T_n, vec[prev] = T_{n-2}
    int t2_rows_normalized = 0;  // launches whose dot rows have been rewritten in the single-step format
    int pair_grid_x = 0;
    PairWalk pair_walk;
    bool cube = false;         // t2 on a three-dimensional lattice: cheb_cube.cu (4-column panels)
    int cube_shape = 0;        // which patch shape of cheb_cube.cu
    CubeWalk cube_walk;
    // dots[(step * 2 + which) * n_panels * PW + panel * PW + c]; which 0 = <T_n,T_n>, 1 = <T_{n+1},T_n>
    DevBuf dots;
    DevBuf partials;  // per-CTA partial dot products of the step in flight
    DevBuf tickets;   // uint32 [n_panels] arrival counters (last CTA reduces)
    DevBuf mu_tmp;    // staging for moment read-out
    DevBuf obs_tmp;   // staging for observables.cu (coefficients, energies, results)
    DevBuf work_items;  // work lists of the two-step kernel (PairWalk)
    int grid_x = 0;
    int panels_per_group = 1;  // ELL kernel: panels sharing one pass over the matrix (grid.y = groups)
    int panel_batch = 1;       // ... of which this many have their loads in flight together
    int n_groups = 0;
    RowWalk walk;              // ELL / DICT kernels: row traversal
    int64_t launches = 0;
};

// Kernel-native copy of the packed matrix for the ELL step kernel (cheb_ell.cu): every block row
// padded to `width` slots, slot 0 = the diagonal block (zero block if absent), then the other
// blocks in ascending column order; padding slots are zero blocks pointing at the row itself.
//   idx  int32  [n_sites][width]
//   data double [n_sites][width][4 (a)][2 (re, im)][4 (b)]   -- MMA B-fragment order, 256 B/block
struct EllDev {
    bool valid = false;
    bool usable = false;   // false: rows too long / too ragged, use the generic kernel
    int width = 0;
    int64_t n_sites = 0;
    DevBuf idx, data;
    // Block dictionary over the same slots (cheb_ell.cu, DICT kernels): the DISTINCT blocks of the
    // matrix, and one code per slot.  Usable when the distinct blocks are a small fraction of all.
    //   code  int32  [n_sites][width]          table double [n_unique][4][2][4]
    bool dict_usable = false;
    bool diag_usable = false;  // every block outside slot 0 is real and diagonal: dtab[n_unique][4] holds the diagonals
    bool self_diag_usable = false;  // every block IN slot 0 (on-site) is real and diagonal (cheb_step_ell<..., SD>)
    int64_t n_unique = 0;
    DevBuf code, table, dtab;
    // Two-steps-per-pass kernel (cheb_pair.cu): dictionary format + nearest-neighbour stencil on a
    // lattice whose x-planes are one-dimensional (pair_M sites per plane).
    bool pair_usable = false;
    bool pair_open = false;  // no block wraps around in-plane
    int pair_M = 0;
    DevBuf dcode;  // int32 [n_sites][5]: dictionary code per stencil direction (self, x-1, y-1, y+1, x+1), -1 = none
    // ... and its three-dimensional sibling (cheb_cube.cu): open nearest-neighbour stencil on a lattice with Ly, Lz >= 2.
    bool cube_usable = false;
    DevBuf dcode3;  // int32 [n_sites][8]: (self, x-1, y-1, z-1, z+1, y+1, x+1, pad); -1 = no block, -2 = no site in y / z
    DevBuf tmp_keys, tmp_rep, tmp_where, tmp_dense;  // hash-table scratch of the dictionary build (kept for rebuilds)
    // ... and what stays alive for incremental updates (ell_patch): the hash table itself (tmp_keys, `hash_cap` slots),
    // posid[hash position] = dictionary code (-1: unused), the table's capacity in entries, device counters
    // {n_unique, pattern changed, table full, hash collision, off-site block not real-diagonal}.
    DevBuf posid, counters;
    int64_t hash_cap = 0, table_cap = 0;
};

struct bdg_system {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    bool quiesced = false;     // being destroyed, stream drained: dev_free need not synchronise per buffer
    int64_t dev_bytes = 0;
    int cubic[3] = {0, 0, 0};  // (Lx, Ly, Lz) when built by bdg_create_cubic, else zeros

    BsrDev skel;           // full skeleton, values as scattered so far
    BsrDev packed;         // after eliminate_zeros (built lazily)
    bool packed_valid = false;
    DevBuf pack_flags;     // int32 [skel.n_blocks]: block kept in `packed` (non-zero)      } kept after the compaction so that a
    DevBuf pack_pos;       // int32 [skel.n_blocks]: its position there                      } later scatter can patch the copies
    // Hermiticity bookkeeping: the stored matrix passed max|M - M^H| <= herm_tol and every change since went through a
    // checked scatter -- then the next scatter only needs to look at the blocks it writes (and their transposes).
    bool herm_verified = false;
    double herm_tol = 0.0;
    int64_t stats[5] = {0, 0, 0, 0, 0};  // bdg_stats
    int packed_max_row = 0;  // longest block row of `packed` (picks the kernel's unroll)

    DevBuf scratch_i32[4];  // reusable scratch (counts, scans, flags, positions)
    DevBuf stage[10];       // uploaded entry lists: {i, j, values, k1, k2} x {hopping, pairing}
    DevBuf scalars;         // small device scalars (first_bad, max_dev, ...)
    void *host_scalars = nullptr;  // pinned mirror

    ChebState cheb;
    EllDev ell;
    DevBuf multi_send, multi_recv;  // moments of this GPU / of all GPUs around the NCCL collective (multi.cu)
};

// scan.cu
int exclusive_scan_i32(bdg_system *sys, const int32_t *in, int32_t *out, int64_t n,
                       int32_t *total_dev /* device, may be null */);

// assemble.cu
int build_packed(bdg_system *sys);
int ensure_scratch(bdg_system *sys, int which, size_t bytes);

// cheb.cu
void cheb_release(bdg_system *sys);     // free all Chebyshev buffers
void cheb_deactivate(bdg_system *sys);  // matrix changed: recursion state is stale, keep buffers
int t2_finish_dots(bdg_system *sys);    // T2 mode: dot rows -> single-step format (call before reading st.dots)

// cheb_ell.cu
int ell_build(bdg_system *sys);                    // (re)build sys->ell from sys->packed when stale
int ell_configure(bdg_system *sys);                // pick panels_per_group / grid for the current ChebState
int ell_launch_step(bdg_system *sys, bool first, const void *x_cur, void *x_io, double *dots_step);
// After a scatter: bring `packed` and the kernel-native copies (fixed-width rows, dictionary, direction codes) up to date
// for the `n` skeleton blocks listed in `klist` (device; < 0 = skip) instead of rebuilding them.  *ok = false: the zero
// pattern changed, the table is full, ...: the caller invalidates the copies and the next recursion rebuilds them.
int ell_patch(bdg_system *sys, int64_t n, const int32_t *klist, bool *ok);
void ell_release(bdg_system *sys);

// cheb_pair.cu
int pair_probe(bdg_system *sys);      // sets sys->ell.pair_usable / pair_M (called by ell_build)
int pair_configure(bdg_system *sys);  // patch / segment plan and grid for the current ChebState
int pair_launch(bdg_system *sys, const void *x_prev, const void *x_cur, void *x_next1, void *x_next2, double *dots_step);
int t2_launch(bdg_system *sys, bool first, const void *x_cur, void *x_io, double *dots_step);
bool pair_streams_onsite(const bdg_system *sys);  // on-site fragments fetched per row (large dictionaries) instead of held

// cheb_cube.cu
int cube_probe(bdg_system *sys);      // sets sys->ell.cube_usable / dcode3 (called by ell_build)
int cube_configure(bdg_system *sys);  // patch shape / segment plan and grid for the current ChebState
int cube_launch(bdg_system *sys, bool first, const void *x_cur, void *x_io, double *dots_step);

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
