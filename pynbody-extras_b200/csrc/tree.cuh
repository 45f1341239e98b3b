// tree.cuh — device-resident octree shared by tree_build.cu and tree_walk.cu.
//
// Topology reproduces the reference's recursive bucket octree (tree.rs:804-864) exactly, including
// its node numbering (creation order), but is built without recursion: float64 octant-path keys by
// the reference's own descent arithmetic, radix sort, level-by-level node emission from sorted key
// ranges, then a renumbering pass.
#pragma once
#include <functional>
#include <memory>

#include "common.cuh"

namespace pnbx {

constexpr int KEY_LEVELS_HI = 21;  // 3 bits per level in a 64-bit word (63 bits used)
constexpr int KEY_LEVELS = 42;     // hi + lo words

// One 64-byte, 64-byte-aligned record per node: everything a visit needs in two 32-byte sectors of one line,
// fetched with a single broadcast load per warp.
struct alignas(64) NodeRec {
    double com[3];
    double size2;         // (2*half)^2, tree.rs:794-798
    double gate2;         // (c * max(hmax,0))^2 for the tree's current kernel (tree.rs:61-70), 0 without hmax
    int32_t next_branch;  // tree.rs:736-776; -1 = end of walk
    int32_t first;        // internal: first_subnode; leaf: first particle (sorted order)
    int32_t kind;         // >= 0: leaf with `kind` particles; -1: internal; -2: zero mass (skip subtree)
    int32_t nleaf;        // leaf records only: number of reference leaves this record stands for (see below)
    int32_t ref;          // the node's id in the reference numbering (the float64 payload arrays are indexed by it)
    int32_t pad_;
};
// Walk records (and the fp32 moment records) are stored in DEPTH-FIRST order, not in the reference numbering: the walk
// is a depth-first sweep with skips, so `first` is the next record in memory and every warp moves monotonically forward
// through the array (the reference numbering keeps siblings together but scatters subtrees). Links are remapped.
// Leaf runs. Every lane that opens a node visits ALL of its children, and leaves are always summed directly, so a
// maximal run of consecutive non-zero-mass sibling leaves (consecutive reference ids, contiguous sorted particles) is
// visited by the same lanes one leaf after the other. The walk records merge such a run into the record of its first
// leaf: `first`/`kind` cover the particles of the whole run, `next_branch` is the last leaf's, `nleaf` the number of
// leaves, and `com` is the run's origin for the fp32 sources (mean of the leaf COMs). The other leaves of the run
// keep their single-leaf records but are unreachable: nothing links to them (links and resume points only ever target
// a first child or the sibling after an internal node). The interaction list of every target is unchanged.
static_assert(sizeof(NodeRec) == 64, "NodeRec must be one 64-byte record");

struct pnbx_tree_impl {
    // All of the tree's memory is allocated, built and freed on `stream`, the library's per-device tree stream (never a
    // caller's). Evaluations may run on any other stream: they wait for `ready` (recorded after every build / setter)
    // and, when they finish, make the tree stream wait for them, so a later setter or the destructor can never free or
    // overwrite what a walk is still reading.
    cudaStream_t stream = nullptr;
    cudaEvent_t ready = nullptr;
    Arena a_src, a_topo, a_payload;  // slabs behind the views below (n-sized, node-sized, payload)
    pnbx_tree_impl() = default;
    pnbx_tree_impl(const pnbx_tree_impl&) = delete;
    pnbx_tree_impl& operator=(const pnbx_tree_impl&) = delete;
    ~pnbx_tree_impl();

    // PNBX_DEVICES: this tree is the copy on multi_devs[0]; replicas[r-1] is the identical tree on multi_devs[r]
    std::vector<std::unique_ptr<pnbx_tree_impl>> replicas;
    std::vector<int> multi_devs;

    int device = 0;
    int64_t n = 0;
    int64_t leaf_capacity = 1;
    int order = 0;   // multipole_order clamped to 5
    int order_raw = 0;
    int kernel = PNBX_KERNEL_PLUMMER;
    bool has_mass = false, has_h = false, has_payload = false, has_hmax = false;

    // sources, original order (owned copies, gravity.rs:154-180)
    DevBuf<double> pos, mass, h;
    // root cube (tree.rs:628-654) on the device: {cx, cy, cz, half}
    DevBuf<double> root4;

    // per particle
    DevBuf<uint64_t> key_hi, key_lo;   // original order
    DevBuf<uint32_t> perm;             // sorted position -> original index (ascending inside a leaf)
    // sorted copies used by payload builds and the walk
    DevBuf<double> spos;               // (n,3) float64
    DevBuf<double> smass, sh;          // float64 (smass only if has_mass, sh only if has_h)
    DevBuf<float4> src32;              // (x, y, z, m) float32, relative to the origin of the particle's leaf run
    DevBuf<float> sh32;                // max(h,0)^2 in float32, sorted order

    // nodes, reference numbering
    int64_t nn = 0, n_leaves = 0;
    int depth = 0;
    DevBuf<double> center, half;       // (nn,3), (nn)
    DevBuf<uint8_t> node_depth, node_nchild;
    DevBuf<uint32_t> node_start, node_count;  // particle range in sorted order (all nodes)
    DevBuf<int32_t> first_subnode, next_branch;
    DevBuf<uint64_t> path_hi, path_lo;
    // INTERNAL nodes grouped by level (reference ids, for the bottom-up payload sweeps): those of level d are
    // level_ids[ilevel_off[d] .. ilevel_off[d+1]); the leaves follow behind all internal nodes
    DevBuf<int32_t> level_ids;
    DevBuf<int32_t> dfs_of_ref;        // depth-first index of every node (the order of the walk records)
    std::vector<int64_t> ilevel_off;
    int64_t n_internal = 0;

    // payloads
    DevBuf<double> nmass, ncom, hmax;  // (nn), (nn,3), (nn)
    int n_moments = 0;
    DevBuf<double> moments;            // (nn, n_moments) float64
    DevBuf<float> moments32;           // fp32 walk records, rec32 floats per node (multipole.cuh: m2p_fast layout for
    int rec32 = 0;                     // order <= 3, the plain coefficients padded to a multiple of 4 for orders 4, 5)
    DevBuf<NodeRec> rec;               // walk records (reference numbering)
};

// tree_build.cu — construction in phases, shared by the single- and multi-device entry points.
// tree_init: device, stream, options and the n-sized slab (the caller then fills t.pos / t.mass / t.h on t.stream);
// tree_build: keys, sort, topology and, iff has_mass, the payloads (Octree::new, gravity.rs:121-226).
void tree_init(pnbx_tree_impl& t, int device, int64_t n, int64_t leaf_capacity, int multipole_order, int kernel,
               bool has_mass, bool has_h);
void tree_build(pnbx_tree_impl& t, StageTimer& tm);
void tree_build_mass(pnbx_tree_impl& t, StageTimer& tm, bool regather);  // build_mass (tree.rs:968-1012) on t.stream;
                                                                         // regather: masses were replaced
void tree_mark_ready(pnbx_tree_impl& t);                              // record `ready` on t.stream
void tree_begin_use(const pnbx_tree_impl& t, cudaStream_t s);         // s waits for the tree to be ready
void tree_end_use(const pnbx_tree_impl& t, cudaStream_t s);           // t.stream waits for the work queued on s
void stream_wait_stream(cudaStream_t waiter, cudaStream_t on, int device);

// tree_walk.cu
void tree_walk(const pnbx_tree_impl& t, const Exec& ex, const double* d_tgt, int64_t m, int64_t tgt_begin,
               double theta, int want, double* d_pot, double* d_acc, StageTimer& tm, unsigned long long* d_counters,
               const OutSlices* slices = nullptr);

// multi.cu
bool multi_tree_create(pnbx_tree_impl& primary, const double* pos, const double* mass, const double* h, int64_t n,
                       int64_t leaf_capacity, int multipole_order, int kernel, const pnbx_opts* opts);
void multi_tree_for_each(pnbx_tree_impl& primary, const std::function<void(pnbx_tree_impl&)>& fn);
bool multi_tree_eval(pnbx_tree_impl& primary, const double* tgt_pos, int64_t m, double theta, int want, double* out_pot,
                     double* out_acc, const pnbx_opts* opts);

}  // namespace pnbx
