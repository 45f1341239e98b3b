/*
 * pnbx_gravity.h — C-ABI of libpnbx_gravity.so, the B200 (sm_100a) gravity hot path.
 *
 * This is the drop-in boundary for pynbody-extras' gravity path. It replaces the
 * PyO3 module `pynbodyext._rust` (reference: crates/pynbodyext-rust/src/lib.rs:10-27,
 * crates/pynbodyext-rust/src/gravity.rs). Every entry point below names the reference
 * interface it stands in for. A maintainer binds it with ctypes (see INTEGRATION.md and
 * pynbody-extras_b200/pynbodyext/_rust.py).
 *
 * Conventions
 *   - plain C: pointers, sizes, ints; no C++/torch types.
 *   - every function returns 0 on success, non-zero on failure; pnbx_last_error()
 *     returns a thread-local, NUL-terminated message for the last failure.
 *   - positions are row-major (N,3) float64, masses / softenings (N,) float64,
 *     outputs (M,) / (M,3) float64, caller-allocated (same layout as the reference's
 *     numpy arrays, gravity.rs:33-65).
 *   - pnbx_opts.mem_space says whether the pointers are host (library stages H2D/D2H on
 *     its own stream, inputs are copied before compute exactly like gravity.rs:154-180)
 *     or device pointers of the selected device (no copies; stream ordered).
 *   - G = 1, like the reference core (units are applied in Python, pyn_gravity.py:121).
 *   - There is NO CPU fallback: without a CUDA device every compute call fails with
 *     PNBX_ERR_CUDA.
 */
#ifndef PNBX_GRAVITY_H
#define PNBX_GRAVITY_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PNBX_ABI_VERSION 1

/* status codes */
#define PNBX_OK 0
#define PNBX_ERR_ARG 1     /* bad argument (maps to Python ValueError)           */
#define PNBX_ERR_CUDA 2    /* CUDA runtime failure / no device (RuntimeError)     */
#define PNBX_ERR_STATE 3   /* e.g. mass payload not built (ValueError)            */
#define PNBX_ERR_DEPTH 4   /* tree deeper than the path-key width (ValueError)    */

/* softening kernel codes: reference KernelKind (kernel.rs:3-12, base.py:71-80).
 * PNBX_KERNEL_NONE selects the Newtonian direct_* functions (gravity.rs:486-497). */
#define PNBX_KERNEL_NONE (-1)
#define PNBX_KERNEL_PLUMMER 0
#define PNBX_KERNEL_SPLINE 1

/* `want` bit mask */
#define PNBX_WANT_POT 1
#define PNBX_WANT_ACC 2

/* pnbx_opts.mem_space */
#define PNBX_MEM_HOST 0
#define PNBX_MEM_DEVICE 1

/* pnbx_opts.precision: arithmetic of the pairwise / multipole interactions.
 * Opening decisions, tree keys, topology and node payloads are always float64. */
#define PNBX_PREC_F32 0
#define PNBX_PREC_F64 1

/* pnbx_opts.flags */
#define PNBX_FLAG_KERNEL_EVENTS 1 /* record CUDA events around the dominant kernel (pnbx_last_kernel_ms) */
#define PNBX_FLAG_TREE_ORDER 2    /* pnbx_tree_eval self mode: tgt_begin/m select SORTED (tree-order) positions and
                                     outputs are in that order; pnbx_tree_get_order gives the original indices.
                                     Tree-order shards keep warps coherent: the multi-GPU sharding of choice. */
#define PNBX_FLAG_BLOCK_CYCLIC 4  /* with PNBX_FLAG_TREE_ORDER: block-cyclic instead of contiguous shards (every rank
                                     samples every region of the tree => equal cost per rank); tgt_begin is ignored,
                                     m must be pnbx_shard_count(n, block, world, rank). */

typedef struct pnbx_opts {
    int32_t device;    /* CUDA ordinal; -1 = current device                       */
    int32_t mem_space; /* PNBX_MEM_HOST | PNBX_MEM_DEVICE                           */
    int32_t precision; /* PNBX_PREC_F32 (default) | PNBX_PREC_F64                   */
    int32_t flags;     /* bit mask of PNBX_FLAG_*                                 */
    void* stream;      /* cudaStream_t to order work on. NULL = a library-owned stream for
                          PNBX_MEM_HOST, the CUDA default stream for PNBX_MEM_DEVICE */
    /* PNBX_FLAG_BLOCK_CYCLIC only: this rank's targets are the tree-order blocks b with b % shard_world ==
     * shard_rank, block = shard_block consecutive tree-order positions (0 = 4096). */
    int32_t shard_rank;
    int32_t shard_world;
    int64_t shard_block;
} pnbx_opts;

typedef struct pnbx_tree pnbx_tree; /* opaque; owns device copies of the sources */

/* Message of the last error on this thread ("" if none). */
const char* pnbx_last_error(void);
int pnbx_abi_version(void);
/* Number of visible CUDA devices (0 without a driver/GPU; never fails). */
int pnbx_device_count(void);

/*
 * Direct summation. Replaces direct_{potentials,accelerations}{,_at_points}_py
 * (gravity.rs:448-709) and through them direct.rs:115-658.
 *
 *   tgt_pos == NULL : self mode on sources [tgt_begin, tgt_begin+m): skip j == i,
 *                     pair softening h = max(h_i, h_j)          (direct.rs:416-437, 494-520)
 *   tgt_pos != NULL : at-points mode, no skip, h = max(h_j, 0)  (direct.rs:566-581, 633-654)
 *   kernel == PNBX_KERNEL_NONE : Newtonian, src_h must be NULL  (gravity.rs:480-484)
 *   src_mass == NULL : unit masses                              (direct.rs:121-128)
 *   want : PNBX_WANT_POT | PNBX_WANT_ACC; the matching out_* must be non-NULL.
 * tgt_begin/m in self mode select a target shard (multi-GPU: one shard per rank).
 */
int pnbx_direct(const double* src_pos, const double* src_mass, const double* src_h, int64_t n,
                const double* tgt_pos, int64_t m, int64_t tgt_begin, int kernel, int want,
                double* out_pot, double* out_acc, const pnbx_opts* opts);

/*
 * Octree. Replaces the `Octree` pyclass (gravity.rs:113-445) and tree.rs.
 * pnbx_tree_create = Octree::new (gravity.rs:121-226): builds topology + walk links
 * (tree.rs:658-776) and, iff mass != NULL, the payloads (tree.rs:968-1012).
 * kernel must be 0/1 here (None maps to Plummer in the Python shim, gravity.rs:77-82).
 */
int pnbx_tree_create(pnbx_tree** out, const double* pos, const double* mass, const double* h,
                     int64_t n, int64_t leaf_capacity, int multipole_order, int kernel,
                     const pnbx_opts* opts);
/* Octree.build_mass (gravity.rs:228-239): optionally replace masses, (re)build payloads. */
int pnbx_tree_build_mass(pnbx_tree* t, const double* mass);
/* Octree.set_softenings (gravity.rs:241-258): setter only, hmax payload NOT rebuilt. */
int pnbx_tree_set_softenings(pnbx_tree* t, const double* h);
/* The same two calls with options: opts->mem_space = PNBX_MEM_DEVICE takes a device pointer of the tree's device,
 * ordered after opts->stream (the caller's stream waits for the update in turn); NULL opts = host pointer. */
int pnbx_tree_build_mass_ex(pnbx_tree* t, const double* mass, const pnbx_opts* opts);
int pnbx_tree_set_softenings_ex(pnbx_tree* t, const double* h, const pnbx_opts* opts);
/* Octree.set_kernel (gravity.rs:260-265). */
int pnbx_tree_set_kernel(pnbx_tree* t, int kernel);
/*
 * compute_{potentials,accelerations} (tgt_pos == NULL; targets = own particles
 * [tgt_begin, tgt_begin+m), skip self, target softening h_i; tree.rs:1415-1496) and
 * {potentials,accelerations}_at_points (tgt_pos != NULL; tree.rs:1498-1558).
 * Fails with PNBX_ERR_STATE "mass payload not built; ..." like gravity.rs:274-278.
 */
int pnbx_tree_eval(pnbx_tree* t, const double* tgt_pos, int64_t m, int64_t tgt_begin, double theta,
                   int want, double* out_pot, double* out_acc, const pnbx_opts* opts);
void pnbx_tree_destroy(pnbx_tree* t);

/* Introspection (used by the parity tests; host pointers only). */
typedef struct pnbx_tree_info {
    int64_t n_particles;
    int64_t n_nodes;
    int64_t n_leaves;
    int32_t depth;           /* deepest node level, root = 0                       */
    int32_t multipole_order; /* as stored (clamped to 5)                           */
    int32_t n_moments;       /* f64 coefficients per node: 0 (no payload) 1/10/20/35/56 */
    int32_t has_payload;
    int32_t has_hmax;
    int32_t kernel;
    int64_t leaf_capacity;
} pnbx_tree_info;
int pnbx_tree_get_info(const pnbx_tree* t, pnbx_tree_info* info);

/*
 * Topology in the reference's node numbering (creation order of tree.rs:804-864):
 *   center[3*i..], half[i], depth[i], first_subnode[i], next_branch[i] (-1 = usize::MAX,
 *   tree.rs:736-776), leaf_start[i]/leaf_count[i] (count -1 for internal nodes;
 *   start indexes `leaf_particles`, which holds every leaf's particle ids in ascending
 *   original order, tree.rs:813-828), path_hi/path_lo: the octant-path key of the node
 *   (3 bits per level, level 1 in the most significant used digit).
 * Any pointer may be NULL to skip that field.
 */
int pnbx_tree_dump_topology(const pnbx_tree* t, double* center, double* half, int32_t* depth,
                            int64_t* first_subnode, int64_t* next_branch, int64_t* leaf_start,
                            int64_t* leaf_count, int64_t* leaf_particles, uint64_t* path_hi,
                            uint64_t* path_lo);
/* Node payloads, float64 as built: mass[i], com[3*i..], hmax[i] (if has_hmax),
 * moments[n_moments*i..] in the reference's field order (multipole.rs:11-74). */
int pnbx_tree_dump_payload(const pnbx_tree* t, double* mass, double* com, double* hmax,
                           double* moments);
/* Per-particle octant-path keys in original particle order (levels 1..21 in hi, 22..42 in lo). */
int pnbx_tree_dump_keys(const pnbx_tree* t, uint64_t* key_hi, uint64_t* key_lo);

/* Traversal statistics for the given targets (same target arguments as pnbx_tree_eval): totals of
 * out5 = {node visits, accepted nodes, leaf visits, leaf particles} summed over targets — the work model's
 * inputs, equal to the reference walk's own counts — and out5[4] = nodes visited per warp summed over warps
 * (the union of 32 targets' paths, i.e. what the warp-cooperative walk really traverses). */
int pnbx_tree_walk_counters(pnbx_tree* t, const double* tgt_pos, int64_t m, int64_t tgt_begin, double theta,
                            int64_t* out5, const pnbx_opts* opts);

/* Number of tree-order positions owned by `rank` under block-cyclic sharding of n positions. */
int64_t pnbx_shard_count(int64_t n, int64_t block, int32_t world, int32_t rank);

/* Original particle index of every tree-order position in [begin, begin+m) (host or device int64 per opts). */
int pnbx_tree_get_order(const pnbx_tree* t, int64_t begin, int64_t m, int64_t* out, const pnbx_opts* opts);
/* (with PNBX_FLAG_BLOCK_CYCLIC in opts->flags: the positions of that rank's block-cyclic shard, `begin` ignored) */

/* Device memory is cached by the library between calls (blocks are reused in stream order, never handed back to the
 * driver behind the caller's back); this returns every cached block that is not in use. */
void pnbx_trim_memory(void);

/* Stage timings of the last call on this thread (GRAVITY_TIMING analogue, tree.rs:5-21):
 * fills up to `cap` (label, milliseconds) pairs, returns the count. */
int pnbx_last_timings(const char** labels, double* ms, int cap);

/* Measurement helpers for bench.py. pnbx_last_kernel_ms: device time of the dominant kernel of the
 * last call made with PNBX_FLAG_KERNEL_EVENTS on this thread (synchronises on its end event).
 * pnbx_launch_count: kernels launched by this library in this process so far. */
int pnbx_last_kernel_ms(double* ms);
int64_t pnbx_launch_count(void);

/* Measurement helper (bench.py roofline denominator): achieved FP32 TFLOP/s of an
 * FMA-chain microbenchmark on `device`. variant 0 = scalar FFMA, 1 = packed fma.rn.f32x2. */
int pnbx_measure_fp32_peak(int device, int variant, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* PNBX_GRAVITY_H */
