// rtb_api.cu — implementation of the C ABI declared in include/rtb.h.
//
// Host-side only: validates and re-lays out the scene for the device (threaded pre-order BVH,
// packed materials/textures), owns device memory behind the opaque handles, launches the kernels
// of rtb_kernels.cu and maps CUDA errors to RtbStatus.  There is NO CPU fallback: without a CUDA
// device every compute entry point fails with RTB_ERR_NO_DEVICE.
#include <algorithm>
#include <cstdlib>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <cuda_fp16.h>

#include "rtb_kernels.cuh"
#include "rtb_wavefront.cuh"

using namespace rtb;

// ------------------------------------------------------------------ errors
static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}
static int cuda_fail(cudaError_t e, const char* what) {
    const int code = (e == cudaErrorMemoryAllocation) ? RTB_ERR_OUT_OF_MEMORY
                     : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? RTB_ERR_NO_DEVICE
                                                                                   : RTB_ERR_CUDA;
    return fail(code, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}
#define RTB_CUDA(call)                                            \
    do {                                                          \
        cudaError_t e__ = (call);                                 \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call);     \
    } while (0)

// ------------------------------------------------------------------ handles
struct RtbScene {
    int device = 0;
    DevScene dev{};
    uint32_t max_depth = 0;  // tree depth (diagnostics)
    std::vector<void*> allocations;
    bool nodes_fit_smem = false;
    unsigned long long* d_counters = nullptr;
    // frame-sized staging buffers of the host-buffer entry points, kept across calls
    float* stage_accum = nullptr;
    uint8_t* stage_rgba = nullptr;
    uint64_t stage_pixels = 0;
    WavefrontState* wavefront = nullptr;
    std::mutex mutex;  // serialises renders on one scene (counters / wavefront queues are shared)
    // The per-octant layouts are built and uploaded on the FIRST use of their traversal mode (ensure_layouts): eight
    // octant copies per mode are ~25x the node memory if all modes are built up front (1.8 GB for the 1 M-sphere scene).
    // What the builders need of the caller's description is kept here (the caller's arrays are not retained).
    std::vector<RtbBvhNode> host_nodes;
    std::vector<RtbHittable> host_hittables;
    std::vector<uint32_t> host_size, host_quad_slot;
    RtbSceneDesc host_desc{};
    bool layout_ready[3] = {false, false, false};  // octant layouts of modes 0 (reference), 1 (ordered), 2 (SAH + SAH16)
};

struct RtbJob {
    std::thread worker;
    std::atomic<uint32_t> samples_done{0};
    uint32_t samples_total = 0;
    std::atomic<int> running{1};
    std::atomic<int> cancel{0};
    int status = RTB_OK;
    std::string error;
    RtbRenderStats stats{};
};

// Library-internal (rtb_multi.cu): sets the calling thread's error message, returns `code`.
extern "C" int rtb_set_last_error(int code, const char* message) {
    g_last_error = message ? message : "";
    return code;
}

extern "C" uint32_t rtb_abi_version(void) { return RTB_ABI_VERSION; }
extern "C" const char* rtb_last_error(void) { return g_last_error.c_str(); }

extern "C" int rtb_device_count(int* count) {
    if (!count) return fail(RTB_ERR_INVALID_ARGUMENT, "rtb_device_count: count is NULL");
    int n = 0;
    const cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        *count = 0;
        cudaGetLastError();
        return fail(RTB_ERR_NO_DEVICE, "no CUDA device visible (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    }
    *count = n;
    return RTB_OK;
}

// ------------------------------------------------------------------ scene lowering
template <class T>
static int upload(RtbScene* sc, const std::vector<T>& host, const T** dev_out) {
    *dev_out = nullptr;
    if (host.empty()) return RTB_OK;
    void* d = nullptr;
    RTB_CUDA(cudaMalloc(&d, host.size() * sizeof(T)));
    sc->allocations.push_back(d);
    RTB_CUDA(cudaMemcpy(d, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    *dev_out = static_cast<const T*>(d);
    return RTB_OK;
}

static inline float4 mkf4(float x, float y, float z, float w) { return make_float4(x, y, z, w); }
static inline float bits(uint32_t u) {
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

// Quad.init's derived fields (src/objects.zig:206-211), evaluated unfused in f32 on the host.
static DevQuad make_quad_quv(const float q[3], const float u[3], const float v[3]);
static DevQuad make_quad(const RtbHittable& h) { return make_quad_quv(h.a, h.b, h.c); }

// Appends the table entries of one complex object and returns its slot: a quad is one entry; a box
// (RTB_HITTABLE_BOX) is a header {offset, sin, cos} followed by createBox's six quads in its order
// (src/objects.zig:510-532 — the z = min face twice, no z = max face, as in the reference).
static uint32_t append_complex(std::vector<DevQuad>& table, const RtbHittable& h) {
    const uint32_t slot = (uint32_t)table.size();
    if (h.type == RTB_HITTABLE_QUAD) {
        table.push_back(make_quad(h));
        return slot;
    }
    DevQuad hdr{};
    hdr.q_d = make_float4(h.c[0], h.c[1], h.c[2], h.sin_theta);
    hdr.u = make_float4(h.cos_theta, h.type == RTB_HITTABLE_CONSTANT_MEDIUM ? h.radius : 0.0f, 0.0f, 0.0f);  // u.y = neg_inv_density
    table.push_back(hdr);
    const float mn[3] = {std::fmin(h.a[0], h.b[0]), std::fmin(h.a[1], h.b[1]), std::fmin(h.a[2], h.b[2])};
    const float mx[3] = {std::fmax(h.a[0], h.b[0]), std::fmax(h.a[1], h.b[1]), std::fmax(h.a[2], h.b[2])};
    const float ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
    const float q[6][3] = {{mn[0], mn[1], mn[2]}, {mx[0], mn[1], mx[2]}, {mx[0], mn[1], mn[2]},
                           {mn[0], mn[1], mn[2]}, {mn[0], mx[1], mx[2]}, {mn[0], mn[1], mn[2]}};
    const float u[6][3] = {{ex, 0, 0}, {-0.0f, -0.0f, -ez}, {-ex, -0.0f, -0.0f}, {0, 0, ez}, {ex, 0, 0}, {ex, 0, 0}};
    const float v[6][3] = {{0, ey, 0}, {0, ey, 0}, {0, ey, 0}, {0, ey, 0}, {-0.0f, -0.0f, -ez}, {0, 0, ez}};
    for (int f = 0; f < 6; ++f) table.push_back(make_quad_quv(q[f], u[f], v[f]));
    return slot;
}

static DevQuad make_quad_quv(const float q[3], const float u[3], const float v[3]) {
    const float n[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
    const float len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    const float normal[3] = {n[0] / len, n[1] / len, n[2] / len};
    const float d = normal[0] * q[0] + normal[1] * q[1] + normal[2] * q[2];
    const float nn = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
    DevQuad dq;
    dq.q_d = mkf4(q[0], q[1], q[2], d);
    dq.u = mkf4(u[0], u[1], u[2], 0);
    dq.v = mkf4(v[0], v[1], v[2], 0);
    dq.normal = mkf4(normal[0], normal[1], normal[2], 0);
    dq.w = mkf4(n[0] / nn, n[1] / nn, n[2] / nn, 0);
    return dq;
}

static int validate_desc(const RtbSceneDesc* d) {
    if (!d) return fail(RTB_ERR_INVALID_ARGUMENT, "scene desc is NULL");
    if (d->abi_version != RTB_ABI_VERSION)
        return fail(RTB_ERR_INVALID_ARGUMENT, "abi_version %u != %u", d->abi_version, RTB_ABI_VERSION);
    if ((d->n_nodes && !d->nodes) || (d->n_hittables && !d->hittables) || (d->n_materials && !d->materials) ||
        (d->n_textures && !d->textures) || (d->n_perlins && !d->perlins) || (d->n_images && !d->images))
        return fail(RTB_ERR_INVALID_ARGUMENT, "scene desc has a NULL array with a non-zero count");
    if (d->n_nodes > RTB_META_INDEX_MASK || d->n_hittables > RTB_META_INDEX_MASK)
        return fail(RTB_ERR_INVALID_ARGUMENT, "scene too large (2^30 nodes/objects max)");
    if (d->n_nodes && (d->root < 0 || (uint32_t)d->root >= d->n_nodes))
        return fail(RTB_ERR_INVALID_ARGUMENT, "root %d out of range", d->root);
    for (uint32_t i = 0; i < d->n_hittables; ++i) {
        const RtbHittable& h = d->hittables[i];
        if (h.type > RTB_HITTABLE_MEDIUM_OF)
            return fail(RTB_ERR_UNSUPPORTED, "hittable %u: unsupported type %u", i, h.type);
        const bool wrapper = h.type == RTB_HITTABLE_TRANSLATE || h.type == RTB_HITTABLE_ROTATE_Y ||
                             h.type == RTB_HITTABLE_LIST || h.type == RTB_HITTABLE_MEDIUM_OF;
        const bool has_material = h.type != RTB_HITTABLE_TRANSLATE && h.type != RTB_HITTABLE_ROTATE_Y && h.type != RTB_HITTABLE_LIST;
        if (has_material && h.material >= d->n_materials)
            return fail(RTB_ERR_INVALID_ARGUMENT, "hittable %u: material out of range", i);
        if (wrapper) {  // children come AFTER their wrapper (no cycles, bounded recursion), list members are contiguous
            const uint64_t count = h.type == RTB_HITTABLE_LIST ? h.material : 1u;
            if (count == 0 || h.child <= i || (uint64_t)h.child + count > d->n_hittables)
                return fail(RTB_ERR_INVALID_ARGUMENT, "hittable %u: child range [%u, +%llu) invalid (children must follow their wrapper)",
                            i, h.child, (unsigned long long)count);
        }
    }
    for (uint32_t i = 0; i < d->n_materials; ++i) {
        const RtbMaterial& m = d->materials[i];
        if (m.type > RTB_MAT_ISOTROPIC) return fail(RTB_ERR_UNSUPPORTED, "material %u: unsupported type %u", i, m.type);
        const bool has_tex = m.type == RTB_MAT_LAMBERTIAN || m.type == RTB_MAT_DIFFUSE_LIGHT || m.type == RTB_MAT_ISOTROPIC;
        if (has_tex && m.texture >= d->n_textures)
            return fail(RTB_ERR_INVALID_ARGUMENT, "material %u: texture out of range", i);
    }
    for (uint32_t i = 0; i < d->n_textures; ++i) {
        const RtbTexture& t = d->textures[i];
        if (t.type > RTB_TEX_NOISE) return fail(RTB_ERR_UNSUPPORTED, "texture %u: unsupported type %u", i, t.type);
        if (t.type == RTB_TEX_IMAGE && t.index >= d->n_images)
            return fail(RTB_ERR_INVALID_ARGUMENT, "texture %u: image out of range", i);
        if (t.type == RTB_TEX_NOISE && t.index >= d->n_perlins)
            return fail(RTB_ERR_INVALID_ARGUMENT, "texture %u: perlin table out of range", i);
    }
    for (uint32_t i = 0; i < d->n_images; ++i) {
        const RtbImage& im = d->images[i];
        if (im.height && im.width && (!im.data || im.bytes_per_row < im.width * 4u))
            return fail(RTB_ERR_INVALID_ARGUMENT, "image %u: bad data/stride", i);
    }
    return RTB_OK;
}

// Pointer graph -> threaded pre-order arrays.  Iterative (no recursion: the caller's tree may be
// arbitrarily deep).  tree_sizes validates the graph and computes subtree sizes in post-order;
// emit_layout assigns pre-order slots: slot(first) = slot + 1, slot(second) = slot + 1 + size(first),
// skip = slot + size(self).
static int tree_sizes(const RtbSceneDesc* d, std::vector<uint32_t>& size, uint32_t* depth_out) {
    *depth_out = 0;
    const uint32_t n = d->n_nodes;
    size.assign(n, 0);
    if (n == 0) return RTB_OK;
    std::vector<uint8_t> state(n, 0);  // 0 = unseen, 1 = expanded, 2 = sized
    std::vector<uint32_t> depth(n, 0);
    std::vector<int32_t> stack;
    stack.push_back(d->root);
    depth[d->root] = 1;
    while (!stack.empty()) {
        const int32_t i = stack.back();
        const RtbBvhNode& nd = d->nodes[i];
        if (depth[i] > *depth_out) *depth_out = depth[i];
        if (nd.leaf >= 0) {
            if ((uint32_t)nd.leaf >= d->n_hittables) return fail(RTB_ERR_INVALID_ARGUMENT, "node %d: leaf out of range", i);
            if (state[i] != 0) return fail(RTB_ERR_INVALID_ARGUMENT, "node %d reached twice (not a tree)", i);
            state[i] = 2;
            size[i] = 1;
            stack.pop_back();
            continue;
        }
        if (nd.left < 0 || nd.right < 0 || (uint32_t)nd.left >= n || (uint32_t)nd.right >= n)
            return fail(RTB_ERR_INVALID_ARGUMENT, "node %d: child out of range", i);
        if (state[i] == 0) {
            state[i] = 1;
            if (state[nd.left] != 0 || state[nd.right] != 0 || nd.left == nd.right)
                return fail(RTB_ERR_INVALID_ARGUMENT, "node %d: child reached twice (not a tree)", i);
            depth[nd.left] = depth[nd.right] = depth[i] + 1;
            stack.push_back(nd.right);
            stack.push_back(nd.left);
        } else {
            size[i] = 1 + size[nd.left] + size[nd.right];
            state[i] = 2;
            stack.pop_back();
        }
    }
    return RTB_OK;
}

static void leaf_record(const RtbSceneDesc* d, uint32_t object, const std::vector<uint32_t>& quad_slot, float4* f0,
                        float4* f1) {
    const RtbHittable& h = d->hittables[object];
    if (h.type == RTB_HITTABLE_SPHERE) {
        const uint32_t kind = h.is_moving ? KIND_MOVING_SPHERE : KIND_SPHERE;
        *f0 = mkf4(h.a[0], h.a[1], h.a[2], bits((kind << 30) | object));
        *f1 = mkf4(h.b[0], h.b[1], h.b[2], h.radius);
    } else if (h.type >= RTB_HITTABLE_TRANSLATE) {  // a general wrapper: the leaf test goes through hit_any(object)
        *f0 = mkf4(0, 0, 0, bits((KIND_QUAD << 30) | object));
        *f1 = mkf4(bits(COMPLEX_GENERIC), 0, 0, bits(object));
    } else {
        *f0 = mkf4(0, 0, 0, bits((KIND_QUAD << 30) | object));
        const uint32_t subtype = h.type == RTB_HITTABLE_BOX ? COMPLEX_BOX
                                 : h.type == RTB_HITTABLE_CONSTANT_MEDIUM ? COMPLEX_MEDIUM : COMPLEX_QUAD;
        *f1 = mkf4(bits(subtype), 0, 0, bits(quad_slot[object]));
    }
}

// octant < 0: bounds as (min, max), reference child order (megakernel / ray queries).
// octant 0..7: bounds pre-swapped to (entry, exit) planes for rays with invD<0 on the axes whose bit
// is set; child order = reference (left first) or, if `ordered`, the child the ray meets first along
// the axis that separates the two children most (the builder split on box-min order, bvh.zig:64-67).
// box_leaves: every leaf is preceded by a box node holding the leaf's own bounds (skip = past the leaf), so the
// primitive test only runs for rays that enter the object's box; a subtree of s nodes then takes s + (s+1)/2 slots.
static inline uint32_t layout_slots(uint32_t subtree_nodes, bool box_leaves) {
    return box_leaves ? subtree_nodes + (subtree_nodes + 1u) / 2u : subtree_nodes;
}
// Near-child-first for an octant: true if rays of this octant meet the RIGHT child first, judged along the axis that
// separates the two children most (the reference's builder split on box-min order, bvh.zig:64-67).
static bool child_order_swapped(const RtbSceneDesc* d, const RtbBvhNode& nd, int octant) {
    const RtbBvhNode& l = d->nodes[nd.left];
    const RtbBvhNode& r = d->nodes[nd.right];
    int axis = 0;
    float best = -1.0f;
    for (int a = 0; a < 3; ++a) {
        const float sep = std::fabs((r.bmin[a] + r.bmax[a]) - (l.bmin[a] + l.bmax[a]));
        if (sep > best) {
            best = sep;
            axis = a;
        }
    }
    const bool left_is_low = (l.bmin[axis] + l.bmax[axis]) <= (r.bmin[axis] + r.bmax[axis]);
    const bool dir_negative = ((octant >> axis) & 1) != 0;
    return left_is_low == dir_negative;
}

static void emit_layout(const RtbSceneDesc* d, const std::vector<uint32_t>& size, const std::vector<uint32_t>& quad_slot,
                        int octant, bool ordered, float4* out, bool box_leaves = false, uint32_t slot_base = 0u) {
    if (d->n_nodes == 0) return;
    struct Item {
        int32_t node;
        uint32_t slot;
    };
    std::vector<Item> work;
    work.push_back({d->root, slot_base});
    if (box_leaves && d->nodes[d->root].leaf < 0) {
        // The root's box test is wasted work (its children's boxes lie inside it and nearly every ray enters it):
        // the SAH layouts start with the root's two subtrees, one slot less.
        work.clear();
        const RtbBvhNode& nd = d->nodes[d->root];
        int32_t first = nd.left, second = nd.right;
        if (ordered && octant >= 0 && child_order_swapped(d, nd, octant)) {
            first = nd.right;
            second = nd.left;
        }
        work.push_back({second, slot_base + layout_slots(size[first], true)});
        work.push_back({first, slot_base});
    }
    while (!work.empty()) {
        const Item it = work.back();
        work.pop_back();
        const RtbBvhNode& nd = d->nodes[it.node];
        if (nd.leaf >= 0 && !box_leaves) {
            leaf_record(d, (uint32_t)nd.leaf, quad_slot, &out[2 * (size_t)it.slot], &out[2 * (size_t)it.slot + 1]);
            continue;
        }
        float lo[3], hi[3];
        for (int a = 0; a < 3; ++a) {
            const bool neg = octant >= 0 && ((octant >> a) & 1);
            lo[a] = neg ? nd.bmax[a] : nd.bmin[a];
            hi[a] = neg ? nd.bmin[a] : nd.bmax[a];
        }
        const uint32_t skip = it.slot + layout_slots(size[it.node], box_leaves);
        out[2 * (size_t)it.slot] = mkf4(lo[0], lo[1], lo[2], bits((KIND_INTERIOR << 30) | skip));
        out[2 * (size_t)it.slot + 1] = mkf4(hi[0], hi[1], hi[2], 0.0f);
        if (nd.leaf >= 0) {  // box_leaves: {box, skip = slot + 2}, {leaf}
            leaf_record(d, (uint32_t)nd.leaf, quad_slot, &out[2 * (size_t)it.slot + 2], &out[2 * (size_t)it.slot + 3]);
            continue;
        }
        int32_t first = nd.left, second = nd.right;
        if (ordered && octant >= 0 && child_order_swapped(d, nd, octant)) {  // the ray comes from the right child's side
            first = nd.right;
            second = nd.left;
        }
        work.push_back({second, it.slot + 1u + layout_slots(size[first], box_leaves)});
        work.push_back({first, it.slot + 1u});
    }
}

// RTB_TRAVERSAL_SAH: re-partition the SAME objects (with the bounding boxes the host computed for them)
// by a binned surface-area-heuristic sweep.  The reference's constructTree picks a random axis and
// splits at the median (src/bvh.zig:48-67), which on Book-1 costs 50 slab tests per ray and on the
// 1 M-sphere scene 774; this tree only changes WHICH nodes are visited — slab test, sphere test and
// every hit value are computed by the same code.  One object per leaf, so it has the same 2n-1 nodes.
// Objects whose box covers at least half of the scene's (the r = 1000 ground sphere of the book scenes) are kept out
// of the tree: `huge` lists them, the layouts test them first, as plain leaves without a box test — every lane of a
// warp does that at the same time (no divergence), their box would be entered by nearly every ray anyway, and the hit
// they usually produce prunes the walk that follows.
static int32_t build_sah(const RtbSceneDesc* d, const std::vector<uint32_t>& reachable, std::vector<RtbBvhNode>& out,
                         std::vector<int32_t>& huge) {
    struct Item {
        float bmin[3], bmax[3], c[3];
        int32_t object;
    };
    std::vector<Item> items;
    items.reserve(d->n_hittables);
    for (uint32_t i = 0; i < d->n_nodes; ++i) {
        const RtbBvhNode& nd = d->nodes[i];
        if (nd.leaf < 0 || reachable[i] == 0) continue;  // only what the host's tree actually references
        Item it;
        for (int a = 0; a < 3; ++a) {
            it.bmin[a] = nd.bmin[a];
            it.bmax[a] = nd.bmax[a];
            it.c[a] = 0.5f * (nd.bmin[a] + nd.bmax[a]);
        }
        it.object = nd.leaf;
        // Quad.init's box is fromPoints(q, q + u + v).pad() (src/objects.zig:209): the two far corners only, which does
        // not cover a quad that is not axis-aligned (q + u and q + v may lie outside it).  The reference never box-tests
        // a leaf (src/bvh.zig:123-125), so there the short box only loosens the parents; this tree DOES box-test its
        // leaves, so the leaf box is widened to all four corners, padded like Aabb.pad (src/aabb.zig:36-43).
        const RtbHittable& h = d->hittables[nd.leaf];
        if (h.type == RTB_HITTABLE_QUAD) {
            for (int a = 0; a < 3; ++a) {
                const float c0 = h.a[a], c1 = h.a[a] + h.b[a], c2 = h.a[a] + h.c[a], c3 = h.a[a] + h.b[a] + h.c[a];
                float lo = std::fmin(std::fmin(c0, c1), std::fmin(c2, c3));
                float hi = std::fmax(std::fmax(c0, c1), std::fmax(c2, c3));
                if (hi - lo < 0.0001f) {  // Interval.expand(delta) of a thin axis
                    lo -= 0.00005f;
                    hi += 0.00005f;
                }
                it.bmin[a] = std::fmin(it.bmin[a], lo);
                it.bmax[a] = std::fmax(it.bmax[a], hi);
                it.c[a] = 0.5f * (it.bmin[a] + it.bmax[a]);
            }
        }
        items.push_back(it);
    }
    // Pad every object box outwards by 2^-21 of the scene's extent on that axis.  The SAH layouts box-test the
    // leaves too and evaluate the slab test with one FMA per plane (slab_miss_fma); for ray origins inside the scene's
    // bounds that differs from the exact (plane - origin) * invD by at most 2^-24 * (2|origin| + |plane|) in position,
    // so with this pad the SAH traversal visits a superset of the leaves an exact test on the host's boxes would.
    {
        float extent[3] = {0.0f, 0.0f, 0.0f};
        for (const Item& it : items)
            for (int a = 0; a < 3; ++a) extent[a] = std::fmax(extent[a], std::fmax(std::fabs(it.bmin[a]), std::fabs(it.bmax[a])));
        for (Item& it : items)
            for (int a = 0; a < 3; ++a) {
                const float pad = std::ldexp(extent[a], -21);
                it.bmin[a] -= pad;
                it.bmax[a] += pad;
            }
    }
    huge.clear();
    auto area = [](const float* mn, const float* mx) {
        const double x = (double)mx[0] - mn[0], y = (double)mx[1] - mn[1], z = (double)mx[2] - mn[2];
        return x * y + y * z + z * x;
    };
#ifndef RTB_SAH_HUGE_FIRST
#define RTB_SAH_HUGE_FIRST 1
#endif
    if (RTB_SAH_HUGE_FIRST && items.size() > 8) {
        float wmn[3] = {INFINITY, INFINITY, INFINITY}, wmx[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (const Item& it : items)
            for (int a = 0; a < 3; ++a) {
                wmn[a] = std::fmin(wmn[a], it.bmin[a]);
                wmx[a] = std::fmax(wmx[a], it.bmax[a]);
            }
        const double whole = area(wmn, wmx);
        std::vector<Item> rest;
        rest.reserve(items.size());
        for (const Item& it : items) {
            if (huge.size() < 4 && area(it.bmin, it.bmax) >= 0.5 * whole) huge.push_back(it.object);
            else rest.push_back(it);
        }
        items.swap(rest);
    }
    out.clear();
    if (items.empty()) return -1;
    out.reserve(2 * items.size());
    struct Task {
        uint32_t lo, hi;
        int32_t parent;
        bool is_right;
    };
    constexpr int kBins = 32;  // measured on Book-1: 16 bins 26.2 slab tests per ray, 32: 25.6, 64 or an exact sweep: 25.7
    constexpr int kMaxBins = kBins;
    std::vector<Task> stack;
    stack.push_back({0u, (uint32_t)items.size(), -1, false});
    while (!stack.empty()) {
        const Task t = stack.back();
        stack.pop_back();
        const int32_t me = (int32_t)out.size();
        RtbBvhNode nd{};
        float cmin[3], cmax[3];
        for (int a = 0; a < 3; ++a) {
            nd.bmin[a] = cmin[a] = INFINITY;
            nd.bmax[a] = cmax[a] = -INFINITY;
        }
        for (uint32_t i = t.lo; i < t.hi; ++i)
            for (int a = 0; a < 3; ++a) {
                nd.bmin[a] = std::fmin(nd.bmin[a], items[i].bmin[a]);
                nd.bmax[a] = std::fmax(nd.bmax[a], items[i].bmax[a]);
                cmin[a] = std::fmin(cmin[a], items[i].c[a]);
                cmax[a] = std::fmax(cmax[a], items[i].c[a]);
            }
        nd.left = nd.right = -1;
        nd.leaf = -1;
        if (t.parent >= 0) {
            if (t.is_right) out[(size_t)t.parent].right = me;
            else            out[(size_t)t.parent].left = me;
        }
        if (t.hi - t.lo == 1) {
            nd.leaf = items[t.lo].object;
            out.push_back(nd);
            continue;
        }
        out.push_back(nd);
        double best_cost = INFINITY;
        int best_axis = -1, best_bin = 0;
        for (int a = 0; a < 3; ++a) {
            const float extent = cmax[a] - cmin[a];
            if (!(extent > 0.0f)) continue;
            uint32_t cnt[kMaxBins] = {0};
            float bmn[kMaxBins][3], bmx[kMaxBins][3];
            for (int b = 0; b < kBins; ++b)
                for (int k = 0; k < 3; ++k) {
                    bmn[b][k] = INFINITY;
                    bmx[b][k] = -INFINITY;
                }
            const float scale = (float)kBins / extent;
            for (uint32_t i = t.lo; i < t.hi; ++i) {
                int b = (int)((items[i].c[a] - cmin[a]) * scale);
                b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
                ++cnt[b];
                for (int k = 0; k < 3; ++k) {
                    bmn[b][k] = std::fmin(bmn[b][k], items[i].bmin[k]);
                    bmx[b][k] = std::fmax(bmx[b][k], items[i].bmax[k]);
                }
            }
            double right_area[kMaxBins];
            uint32_t right_cnt[kMaxBins];
            float rmn[3] = {INFINITY, INFINITY, INFINITY}, rmx[3] = {-INFINITY, -INFINITY, -INFINITY};
            uint32_t rc = 0;
            for (int b = kBins - 1; b > 0; --b) {
                for (int k = 0; k < 3; ++k) {
                    rmn[k] = std::fmin(rmn[k], bmn[b][k]);
                    rmx[k] = std::fmax(rmx[k], bmx[b][k]);
                }
                rc += cnt[b];
                right_cnt[b] = rc;
                right_area[b] = rc ? area(rmn, rmx) : 0.0;
            }
            float lmn[3] = {INFINITY, INFINITY, INFINITY}, lmx[3] = {-INFINITY, -INFINITY, -INFINITY};
            uint32_t lc = 0;
            for (int b = 0; b < kBins - 1; ++b) {
                for (int k = 0; k < 3; ++k) {
                    lmn[k] = std::fmin(lmn[k], bmn[b][k]);
                    lmx[k] = std::fmax(lmx[k], bmx[b][k]);
                }
                lc += cnt[b];
                if (lc == 0 || right_cnt[b + 1] == 0) continue;
                const double cost = area(lmn, lmx) * lc + right_area[b + 1] * right_cnt[b + 1];
                if (cost < best_cost) {
                    best_cost = cost;
                    best_axis = a;
                    best_bin = b;
                }
            }
        }
        uint32_t mid = t.lo + (t.hi - t.lo) / 2;
        if (best_axis >= 0) {
            const float scale = (float)kBins / (cmax[best_axis] - cmin[best_axis]);
            const float c0 = cmin[best_axis];
            const int a = best_axis, split = best_bin;
            auto* first = items.data() + t.lo;
            auto* last = items.data() + t.hi;
            auto* m = std::partition(first, last, [=](const Item& it) {
                int b = (int)((it.c[a] - c0) * scale);
                b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
                return b <= split;
            });
            const uint32_t pm = (uint32_t)(m - items.data());
            if (pm > t.lo && pm < t.hi) mid = pm;
        }
        stack.push_back({mid, t.hi, me, true});
        stack.push_back({t.lo, mid, me, false});
    }
    return 0;
}

// The RTB_TRAVERSAL_SAH layout of one octant: the huge objects as leading leaves, then the tree of the rest (boxed
// leaves, no root box).  SahTree is built once per scene.
struct SahTree {
    std::vector<RtbBvhNode> nodes;
    std::vector<uint32_t> size;
    std::vector<int32_t> huge;
    RtbSceneDesc desc{};
    uint32_t layout_nodes = 0;  // entries per octant, without the end sentinel
};
static int sah_tree_build(const RtbSceneDesc* desc, const std::vector<uint32_t>& reachable, SahTree& t) {
    t.desc = *desc;
    t.desc.root = build_sah(desc, reachable, t.nodes, t.huge);
    t.desc.nodes = t.nodes.data();
    t.desc.n_nodes = (uint32_t)t.nodes.size();
    uint32_t depth = 0;
    if (t.desc.n_nodes) {
        const int rc = tree_sizes(&t.desc, t.size, &depth);
        if (rc != RTB_OK) return rc;
    }
    const uint32_t rest = t.desc.n_nodes ? t.size[t.desc.root] : 0u;
    t.layout_nodes = (uint32_t)t.huge.size() + (rest ? layout_slots(rest, true) - (rest > 1 ? 1u : 0u) : 0u);
    return RTB_OK;
}
static void sah_emit(const RtbSceneDesc* desc, const SahTree& t, const std::vector<uint32_t>& quad_slot, int octant,
                     float4* out) {
    for (size_t h = 0; h < t.huge.size(); ++h) leaf_record(desc, (uint32_t)t.huge[h], quad_slot, &out[2 * h], &out[2 * h + 1]);
    emit_layout(&t.desc, t.size, quad_slot, octant, true, out, true, (uint32_t)t.huge.size());
}

// RTB_TRAVERSAL_SAH16: the 8 SAH layouts (oct8 = [octant][2 * (n_entries + 1)] float4, as built by sah_emit) packed into
// 16-byte slots — see DevScene::pk_nodes.  Box planes go to the normalised frame n = (x - center) / scale of the box
// nodes' own bounds and are rounded OUTWARDS to binary16 (entry planes away from the box along the ray's direction of
// travel, exit planes likewise), so a packed box always contains the f32 box it was made from.
struct PackedLayout {
    std::vector<uint4> slots;  // [octant][n_slots]
    uint32_t n_slots = 0;      // per octant, sentinel included
    float center[3] = {0, 0, 0}, scale[3] = {1, 1, 1};
};
static inline uint32_t fbits(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
static inline uint32_t half_bits(__half h) {
    uint16_t u;
    std::memcpy(&u, &h, 2);
    return u;
}
static void pack_layout(const std::vector<float4>& oct8, uint32_t n_entries, PackedLayout& out) {
    const size_t stride = 2 * ((size_t)n_entries + 1);
    // frame: bounds of the box nodes (octant 0 holds the same set of boxes as every other octant)
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t k = 0; k < n_entries; ++k) {
        const float4 f0 = oct8[2 * (size_t)k], f1 = oct8[2 * (size_t)k + 1];
        if (fbits(f0.w) >= (1u << 30)) continue;
        const float lo[3] = {f0.x, f0.y, f0.z}, hi[3] = {f1.x, f1.y, f1.z};
        for (int a = 0; a < 3; ++a) {
            mn[a] = std::fmin(mn[a], std::fmin(lo[a], hi[a]));
            mx[a] = std::fmax(mx[a], std::fmax(lo[a], hi[a]));
        }
    }
    for (int a = 0; a < 3; ++a) {
        if (!(mn[a] <= mx[a])) mn[a] = mx[a] = 0.0f;  // no box nodes at all
        out.center[a] = 0.5f * (mn[a] + mx[a]);
        const float half = std::fmax(mx[a] - out.center[a], out.center[a] - mn[a]);
        out.scale[a] = half > 0.0f ? half * 1.001f : 1.0f;
    }
    // slot index of every entry (box 1, leaf 2, sentinel 1); identical for all octants (same tree shape per octant?
    // no — child order differs per octant, so the prefix is computed per octant)
    out.n_slots = 0;
    std::vector<uint32_t> slot_of(n_entries + 2);
    for (int oct = 0; oct < 8; ++oct) {
        const float4* L = oct8.data() + (size_t)oct * stride;
        uint32_t acc = 0;
        for (uint32_t k = 0; k <= n_entries; ++k) {
            slot_of[k] = acc;
            const uint32_t meta = fbits(L[2 * (size_t)k].w);
            acc += (meta < (1u << 30) || meta == RTB_META_END) ? 1u : 2u;
        }
        if (oct == 0) {
            out.n_slots = acc;
            out.slots.assign(8 * (size_t)acc, make_uint4(0u, 0u, 0u, RTB_META_END));
        }
        uint4* S = out.slots.data() + (size_t)oct * out.n_slots;
        for (uint32_t k = 0; k <= n_entries; ++k) {
            const float4 f0 = L[2 * (size_t)k], f1 = L[2 * (size_t)k + 1];
            const uint32_t meta = fbits(f0.w);
            uint4& s0 = S[slot_of[k]];
            if (meta == RTB_META_END) {
                s0 = make_uint4(0u, 0u, 0u, RTB_META_END);
            } else if (meta < (1u << 30)) {
                const float e[3] = {f0.x, f0.y, f0.z}, x[3] = {f1.x, f1.y, f1.z};
                uint32_t w[3];
                for (int a = 0; a < 3; ++a) {
                    const bool neg = ((oct >> a) & 1) != 0;  // entry = max plane, exit = min plane
                    const double ne = ((double)e[a] - out.center[a]) / out.scale[a];
                    const double nx = ((double)x[a] - out.center[a]) / out.scale[a];
                    // outwards: the entry plane moves against the direction of travel, the exit plane along it
                    const __half he = neg ? __float2half_ru((float)std::nextafter((float)ne, INFINITY))
                                          : __float2half_rd((float)std::nextafter((float)ne, -INFINITY));
                    const __half hx = neg ? __float2half_rd((float)std::nextafter((float)nx, -INFINITY))
                                          : __float2half_ru((float)std::nextafter((float)nx, INFINITY));
                    w[a] = half_bits(he) | (half_bits(hx) << 16);
                }
                s0 = make_uint4(w[0], w[1], w[2], slot_of[meta]);  // skip link: entry index -> slot index
            } else {
                uint4& s1 = S[slot_of[k] + 1];
                s0 = make_uint4(fbits(f0.x), fbits(f0.y), fbits(f0.z), meta);
                if ((meta >> 30) == KIND_QUAD) s1 = make_uint4(fbits(f1.x), 0u, 0u, fbits(f1.w) | 0x80000000u);
                else                            s1 = make_uint4(fbits(f1.x), fbits(f1.y), fbits(f1.z), fbits(-std::fabs(f1.w)) | 0x80000000u);
            }
        }
    }
}

// The packed SAH16 layout is also built for scenes whose layout does not fit shared memory; it is then walked from
// global memory: half the bytes and one load instead of two per visit against +8 % visits (binary16 planes over the
// whole scene's extent inflate small boxes) = +15 % on the 1 M-sphere scene (profiles/r3h_million_ab.log).
// RTB_PACK_LARGE=0 switches it off (such a scene then renders SAH16 as SAH, as before r3h).
static bool pack_large_scenes() {
    static const bool v = [] { const char* e = std::getenv("RTB_PACK_LARGE"); return !(e && e[0] == '0'); }();
    return v;
}
static int ensure_layouts(RtbScene* sc, uint32_t mode);

static void scene_free(RtbScene* sc) {
    if (!sc) return;
    cudaSetDevice(sc->device);
    if (sc->wavefront) wavefront_destroy(sc->wavefront);
    for (void* p : sc->allocations) cudaFree(p);
    cudaFree(sc->stage_accum);
    cudaFree(sc->stage_rgba);
    delete sc;
}

extern "C" int rtb_scene_create(const RtbSceneDesc* desc, int device, RtbScene** scene_out) {
    if (!scene_out) return fail(RTB_ERR_INVALID_ARGUMENT, "scene_out is NULL");
    *scene_out = nullptr;
    int rc = validate_desc(desc);
    if (rc != RTB_OK) return rc;
    int ndev = 0;
    rc = rtb_device_count(&ndev);
    if (rc != RTB_OK) return rc;
    if (device < 0 || device >= ndev) return fail(RTB_ERR_INVALID_ARGUMENT, "device %d out of range (0..%d)", device, ndev - 1);
    RTB_CUDA(cudaSetDevice(device));

    std::vector<uint32_t> size;
    uint32_t depth = 0;
    rc = tree_sizes(desc, size, &depth);
    if (rc != RTB_OK) return rc;
    std::vector<DevQuad> quads;
    std::vector<uint32_t> quad_slot(desc->n_hittables, 0);
    for (uint32_t i = 0; i < desc->n_hittables; ++i) {
        const uint32_t ty = desc->hittables[i].type;
        if (ty == RTB_HITTABLE_QUAD || ty == RTB_HITTABLE_BOX || ty == RTB_HITTABLE_CONSTANT_MEDIUM)
            quad_slot[i] = append_complex(quads, desc->hittables[i]);
    }
    const uint32_t n_tree = desc->n_nodes ? size[desc->root] : 0u;
    std::vector<float4> nodes(2 * (size_t)n_tree);
    emit_layout(desc, size, quad_slot, -1, false, nodes.data());
    RtbScene* sc = new (std::nothrow) RtbScene();
    if (!sc) return fail(RTB_ERR_OUT_OF_MEMORY, "host allocation failed");
    sc->device = device;
    sc->max_depth = depth;

    auto has_material = [&](uint32_t i) {
        const uint32_t ty = desc->hittables[i].type;
        return ty != RTB_HITTABLE_TRANSLATE && ty != RTB_HITTABLE_ROTATE_Y && ty != RTB_HITTABLE_LIST;
    };
    std::vector<uint32_t> obj_mat(desc->n_hittables);
    for (uint32_t i = 0; i < desc->n_hittables; ++i) obj_mat[i] = has_material(i) ? desc->hittables[i].material : 0u;

    std::vector<float4> mats(2 * (size_t)desc->n_materials);
    for (uint32_t i = 0; i < desc->n_materials; ++i) {
        const RtbMaterial& m = desc->materials[i];
        const bool has_tex = m.type == RTB_MAT_LAMBERTIAN || m.type == RTB_MAT_DIFFUSE_LIGHT || m.type == RTB_MAT_ISOTROPIC;
        uint32_t tex_type = RTB_TEX_SOLID;
        float rgb[3] = {m.albedo[0], m.albedo[1], m.albedo[2]};
        if (has_tex) {
            const RtbTexture& t = desc->textures[m.texture];
            tex_type = t.type;
            if (t.type == RTB_TEX_SOLID) std::memcpy(rgb, t.color, sizeof(rgb));  // inline the solid colour
        }
        mats[2 * (size_t)i] = mkf4(bits(m.type | (tex_type << 8)), bits(has_tex ? m.texture : 0u), m.fuzz, m.ir);
        mats[2 * (size_t)i + 1] = mkf4(rgb[0], rgb[1], rgb[2], 0.0f);
    }
    // per-object record for the wavefront shader: leaf record + the object's material record, and its class
    std::vector<float4> prims(4 * (size_t)desc->n_hittables);
    std::vector<uint8_t> obj_class(desc->n_hittables);
    for (uint32_t i = 0; i < desc->n_hittables; ++i) {
        leaf_record(desc, i, quad_slot, &prims[4 * (size_t)i], &prims[4 * (size_t)i + 1]);
        obj_class[i] = CLASS_OTHER;
        if (!has_material(i) || desc->n_materials == 0) continue;  // a wrapper: shaded with the inner primitive's material
        const uint32_t m = desc->hittables[i].material;
        prims[4 * (size_t)i + 2] = mats[2 * (size_t)m];
        prims[4 * (size_t)i + 3] = mats[2 * (size_t)m + 1];
        const RtbMaterial& mat = desc->materials[m];
        uint8_t cls = CLASS_OTHER;
        if (mat.type == RTB_MAT_LAMBERTIAN && desc->textures[mat.texture].type == RTB_TEX_SOLID) cls = CLASS_LAMBERT_SOLID;
        if (mat.type == RTB_MAT_METAL) cls = CLASS_METAL;
        if (mat.type == RTB_MAT_DIELECTRIC) cls = CLASS_DIELECTRIC;
        obj_class[i] = cls;
    }
    std::vector<float4> texs(3 * (size_t)desc->n_textures);
    for (uint32_t i = 0; i < desc->n_textures; ++i) {
        const RtbTexture& t = desc->textures[i];
        texs[3 * (size_t)i] = mkf4(bits(t.type), bits(t.index), t.scale, 0.0f);
        texs[3 * (size_t)i + 1] = mkf4(t.color[0], t.color[1], t.color[2], 0.0f);
        texs[3 * (size_t)i + 2] = mkf4(t.color2[0], t.color2[1], t.color2[2], 0.0f);
    }
    std::vector<DevPerlin> perlins(desc->n_perlins);
    for (uint32_t i = 0; i < desc->n_perlins; ++i) {
        const RtbPerlin& p = desc->perlins[i];
        for (int k = 0; k < 256; ++k) {
            perlins[i].ranvec[k] = mkf4(p.ranvec[k][0], p.ranvec[k][1], p.ranvec[k][2], 0.0f);
            perlins[i].perm_x[k] = (uint8_t)(p.perm_x[k] & 255u);
            perlins[i].perm_y[k] = (uint8_t)(p.perm_y[k] & 255u);
            perlins[i].perm_z[k] = (uint8_t)(p.perm_z[k] & 255u);
        }
    }
    std::vector<DevImage> images(desc->n_images);
    for (uint32_t i = 0; i < desc->n_images && rc == RTB_OK; ++i) {
        const RtbImage& im = desc->images[i];
        images[i].width = im.width;
        images[i].height = im.height;
        images[i].texels = nullptr;
        if (im.width && im.height) {
            std::vector<uchar4> tight((size_t)im.width * im.height);
            for (uint32_t y = 0; y < im.height; ++y)
                std::memcpy(&tight[(size_t)y * im.width], im.data + (size_t)y * im.bytes_per_row, (size_t)im.width * 4);
            rc = upload(sc, tight, &images[i].texels);
        }
    }
    if (rc == RTB_OK) rc = upload(sc, nodes, &sc->dev.nodes);
    // keep what ensure_layouts needs (nodes + hittables of the description, subtree sizes, quad-table slots)
    sc->host_nodes.assign(desc->nodes, desc->nodes + desc->n_nodes);
    sc->host_hittables.assign(desc->hittables, desc->hittables + desc->n_hittables);
    sc->host_size = size;
    sc->host_quad_slot = quad_slot;
    sc->host_desc = *desc;
    sc->host_desc.nodes = sc->host_nodes.data();
    sc->host_desc.hittables = sc->host_hittables.data();
    sc->host_desc.materials = nullptr;  // not needed by the layout builders, and not retained
    sc->host_desc.textures = nullptr;
    sc->host_desc.perlins = nullptr;
    sc->host_desc.images = nullptr;
    for (int a = 0; a < 3; ++a) {
        sc->dev.pk_center[a] = 0.0f;
        sc->dev.pk_scale[a] = sc->dev.pk_inv_scale[a] = 1.0f;
    }
    if (rc == RTB_OK) rc = upload(sc, prims, &sc->dev.prims);
    if (rc == RTB_OK) rc = upload(sc, obj_class, &sc->dev.object_class);
    if (rc == RTB_OK) rc = upload(sc, obj_mat, &sc->dev.object_material);
    if (rc == RTB_OK) rc = upload(sc, mats, &sc->dev.materials);
    if (rc == RTB_OK) rc = upload(sc, texs, &sc->dev.textures);
    if (rc == RTB_OK) rc = upload(sc, perlins, &sc->dev.perlins);
    if (rc == RTB_OK) rc = upload(sc, images, &sc->dev.images);
    std::vector<DevInst> insts(desc->n_hittables);
    for (uint32_t i = 0; i < desc->n_hittables; ++i) {
        const RtbHittable& h = desc->hittables[i];
        DevInst in{};
        in.type = h.type;
        in.child = h.child;
        in.count = h.type == RTB_HITTABLE_LIST ? h.material : 1u;
        in.slot = quad_slot[i];
        if (h.type == RTB_HITTABLE_TRANSLATE) in.p = mkf4(h.a[0], h.a[1], h.a[2], 0.0f);
        else if (h.type == RTB_HITTABLE_ROTATE_Y) in.p = mkf4(h.sin_theta, h.cos_theta, 0.0f, 0.0f);
        else if (h.type == RTB_HITTABLE_MEDIUM_OF) in.p = mkf4(h.radius, 0.0f, 0.0f, 0.0f);
        else in.p = mkf4(0.0f, 0.0f, 0.0f, 0.0f);
        insts[i] = in;
    }
    if (rc == RTB_OK) rc = upload(sc, quads, &sc->dev.quads);
    if (rc == RTB_OK) rc = upload(sc, insts, &sc->dev.insts);
    if (rc == RTB_OK) {
        void* c = nullptr;
        cudaError_t e = cudaMalloc(&c, 4 * sizeof(unsigned long long));
        if (e != cudaSuccess)
            rc = cuda_fail(e, "cudaMalloc(counters)");
        else {
            sc->allocations.push_back(c);
            sc->d_counters = static_cast<unsigned long long*>(c);
        }
    }
    if (rc != RTB_OK) {
        scene_free(sc);
        return rc;
    }
    sc->dev.n_nodes = (uint32_t)(nodes.size() / 2);
    sc->dev.oct_n_nodes[0] = sc->dev.oct_n_nodes[1] = sc->dev.n_nodes;
    sc->dev.oct_n_nodes[2] = 0u;
    sc->dev.n_objects = desc->n_hittables;
    {   // hit_any recurses once per wrapper level: make sure the device stack holds the deepest chain of this scene
        std::vector<uint32_t> depth(desc->n_hittables, 1u);
        uint32_t deepest = 1u;
        for (uint32_t i = desc->n_hittables; i-- > 0;) {  // children have larger indices than their wrapper
            const RtbHittable& h = desc->hittables[i];
            if (h.type < RTB_HITTABLE_TRANSLATE) continue;
            const uint32_t count = h.type == RTB_HITTABLE_LIST ? h.material : 1u;
            uint32_t d = 0u;
            for (uint32_t k = 0; k < count; ++k) d = std::max(d, depth[h.child + k]);
            depth[i] = d + 1u;
            deepest = std::max(deepest, depth[i]);
        }
        if (deepest > 1u) {
            size_t have = 0;
            // (generous: the kernels that walk a layout in global memory run at 32 registers and spill more per frame)
            const size_t want = 2048u + 1024u * (size_t)deepest;
            if (cudaDeviceGetLimit(&have, cudaLimitStackSize) == cudaSuccess && have < want) {
                const cudaError_t e = cudaDeviceSetLimit(cudaLimitStackSize, want);
                if (e != cudaSuccess) {
                    scene_free(sc);
                    return cuda_fail(e, "cudaDeviceSetLimit(stack size for nested instances)");
                }
            }
        }
    }
    sc->dev.has_quads = 0u;  // any object that is not a plain sphere selects the kernels that can test complex leaves
    for (uint32_t i = 0; i < desc->n_hittables; ++i)
        if (desc->hittables[i].type != RTB_HITTABLE_SPHERE) sc->dev.has_quads = 1u;
    sc->dev.n_perlins = desc->n_perlins;
    sc->nodes_fit_smem = sc->dev.n_nodes > 0 && (size_t)sc->dev.n_nodes * 32u <= megakernel_max_smem_nodes_bytes();
    if (const char* eager = std::getenv("RTB_EAGER_LAYOUTS")) {  // measurement aid: build every mode's layouts now
        if (eager[0] == '1')
            for (uint32_t m = 0; m < 3 && rc == RTB_OK; ++m) rc = ensure_layouts(sc, m);
        if (rc != RTB_OK) {
            scene_free(sc);
            return rc;
        }
    }
    *scene_out = sc;
    return RTB_OK;
}

// Builds and uploads the per-octant layouts of one traversal mode on first use (caller holds scene->mutex and has made
// the scene's device current).  mode 3 (SAH16) shares mode 2's tree: both are built together.
static int ensure_layouts(RtbScene* sc, uint32_t mode) {
    const uint32_t slot = mode == 3u ? 2u : mode;
    if (slot > 2u || sc->layout_ready[slot]) return RTB_OK;
    const RtbSceneDesc* desc = &sc->host_desc;
    const uint32_t n_tree = sc->dev.n_nodes;
    const float4 sentinel = mkf4(0.0f, 0.0f, 0.0f, bits(RTB_META_END));
    int rc = RTB_OK;
    if (slot < 2u) {
        const size_t oct_stride = 2 * ((size_t)n_tree + 1);
        std::vector<float4> oct(8 * oct_stride, sentinel);
        for (int o = 0; o < 8; ++o)
            emit_layout(desc, sc->host_size, sc->host_quad_slot, o, slot == 1u, oct.data() + (size_t)o * oct_stride);
        rc = upload(sc, oct, &sc->dev.oct_nodes[slot]);
    } else {
        SahTree sah;
        rc = sah_tree_build(desc, sc->host_size, sah);
        if (rc != RTB_OK) return rc;
        const uint32_t n_sah = sah.layout_nodes;
        const size_t sah_stride = 2 * ((size_t)n_sah + 1);
        std::vector<float4> oct(8 * sah_stride, sentinel);
        for (int o = 0; o < 8; ++o) sah_emit(desc, sah, sc->host_quad_slot, o, oct.data() + (size_t)o * sah_stride);
        rc = upload(sc, oct, &sc->dev.oct_nodes[2]);
        sc->dev.oct_n_nodes[2] = n_sah;
        // SAH16: the same layouts packed (walked from shared memory when one octant fits, else from global memory)
        const bool pack_large = pack_large_scenes();
        if (rc == RTB_OK && n_sah > 0 && (pack_large || ((size_t)n_sah + 1) * 32u <= 2 * megakernel_max_smem_nodes_bytes())) {
            PackedLayout packed;
            pack_layout(oct, n_sah, packed);
            if (pack_large || (size_t)packed.n_slots * 16u <= megakernel_max_smem_nodes_bytes()) {
                rc = upload(sc, packed.slots, &sc->dev.pk_nodes);
                sc->dev.pk_slots = packed.n_slots;
                for (int a = 0; a < 3; ++a) {
                    sc->dev.pk_center[a] = packed.center[a];
                    sc->dev.pk_scale[a] = packed.scale[a];
                    sc->dev.pk_inv_scale[a] = 1.0f / packed.scale[a];
                }
            }
        }
    }
    if (rc == RTB_OK) sc->layout_ready[slot] = true;
    return rc;
}

// Test hook: the host-side re-layout of one (mode, octant) without touching a device, so that the CPU
// test-suite can check the threaded layouts (skip links, sentinel, every object exactly once, boxes).
extern "C" int rtb_debug_build_layout(const RtbSceneDesc* desc, uint32_t mode, uint32_t octant, float* out_nodes,
                                      uint32_t* n_nodes_out) {
    int rc = validate_desc(desc);
    if (rc != RTB_OK) return rc;
    if (mode > RTB_TRAVERSAL_SAH || octant > 7 || !n_nodes_out) return fail(RTB_ERR_INVALID_ARGUMENT, "bad mode/octant");
    std::vector<uint32_t> size;
    uint32_t depth = 0;
    rc = tree_sizes(desc, size, &depth);
    if (rc != RTB_OK) return rc;
    std::vector<uint32_t> quad_slot(desc->n_hittables, 0);
    std::vector<DevQuad> table;
    for (uint32_t i = 0; i < desc->n_hittables; ++i) {
        const uint32_t ty = desc->hittables[i].type;
        if (ty == RTB_HITTABLE_QUAD || ty == RTB_HITTABLE_BOX || ty == RTB_HITTABLE_CONSTANT_MEDIUM)
            quad_slot[i] = append_complex(table, desc->hittables[i]);
    }
    const uint32_t n_host = desc->n_nodes ? size[desc->root] : 0u;
    SahTree sah;
    if (mode == RTB_TRAVERSAL_SAH) {
        rc = sah_tree_build(desc, size, sah);
        if (rc != RTB_OK) return rc;
    }
    const uint32_t n_tree = mode == RTB_TRAVERSAL_SAH ? sah.layout_nodes : n_host;
    *n_nodes_out = n_tree;
    if (!out_nodes) return RTB_OK;
    std::vector<float4> layout(2 * ((size_t)n_tree + 1), mkf4(0.0f, 0.0f, 0.0f, bits(RTB_META_END)));
    if (mode == RTB_TRAVERSAL_SAH) {
        sah_emit(desc, sah, quad_slot, (int)octant, layout.data());
    } else {
        emit_layout(desc, size, quad_slot, (int)octant, mode == RTB_TRAVERSAL_ORDERED, layout.data());
    }
    std::memcpy(out_nodes, layout.data(), layout.size() * sizeof(float4));
    return RTB_OK;
}

extern "C" int rtb_debug_packed_layout(const RtbSceneDesc* desc, uint32_t octant, uint32_t* out_slots,
                                       uint32_t* n_slots_out, float center_out[3], float scale_out[3]) {
    int rc = validate_desc(desc);
    if (rc != RTB_OK) return rc;
    if (octant > 7 || !n_slots_out) return fail(RTB_ERR_INVALID_ARGUMENT, "bad octant / n_slots_out");
    std::vector<uint32_t> size;
    uint32_t depth = 0;
    rc = tree_sizes(desc, size, &depth);
    if (rc != RTB_OK) return rc;
    std::vector<uint32_t> quad_slot(desc->n_hittables, 0);
    std::vector<DevQuad> table;
    for (uint32_t i = 0; i < desc->n_hittables; ++i) {
        const uint32_t ty = desc->hittables[i].type;
        if (ty == RTB_HITTABLE_QUAD || ty == RTB_HITTABLE_BOX || ty == RTB_HITTABLE_CONSTANT_MEDIUM)
            quad_slot[i] = append_complex(table, desc->hittables[i]);
    }
    SahTree sah;
    rc = sah_tree_build(desc, size, sah);
    if (rc != RTB_OK) return rc;
    const uint32_t n = sah.layout_nodes;
    if (n == 0 || (!pack_large_scenes() && ((size_t)n + 1) * 32u > 2 * megakernel_max_smem_nodes_bytes()))
        return fail(RTB_ERR_UNSUPPORTED, "scene too large for the packed layout");
    const size_t stride = 2 * ((size_t)n + 1);
    std::vector<float4> oct8(8 * stride, mkf4(0.0f, 0.0f, 0.0f, bits(RTB_META_END)));
    for (int oct = 0; oct < 8; ++oct) sah_emit(desc, sah, quad_slot, oct, oct8.data() + (size_t)oct * stride);
    PackedLayout packed;
    pack_layout(oct8, n, packed);
    if (!pack_large_scenes() && (size_t)packed.n_slots * 16u > megakernel_max_smem_nodes_bytes())
        return fail(RTB_ERR_UNSUPPORTED, "scene too large for the packed layout");
    *n_slots_out = packed.n_slots;
    for (int a = 0; a < 3; ++a) {
        if (center_out) center_out[a] = packed.center[a];
        if (scale_out) scale_out[a] = packed.scale[a];
    }
    if (out_slots) std::memcpy(out_slots, packed.slots.data() + (size_t)octant * packed.n_slots, (size_t)packed.n_slots * 16u);
    return RTB_OK;
}

extern "C" int rtb_scene_destroy(RtbScene* scene) {
    if (!scene) return RTB_OK;
    // A render in flight on this scene (rtb_render_async's worker, or another thread's rtb_render) holds the mutex:
    // wait for it instead of freeing the queues under it.  Jobs must still be waited for / destroyed by the caller
    // before the scene is destroyed if they have batches left to start (rtb.h).
    { std::lock_guard<std::mutex> lock(scene->mutex); }
    scene_free(scene);
    return RTB_OK;
}

// ------------------------------------------------------------------ ray queries
extern "C" int rtb_trace_rays(RtbScene* scene, const RtbRay* rays, uint64_t n, uint32_t traversal, RtbHit* hits_out) {
    if (!scene) return fail(RTB_ERR_INVALID_ARGUMENT, "scene is NULL");
    if (n && (!rays || !hits_out)) return fail(RTB_ERR_INVALID_ARGUMENT, "rays/hits_out is NULL");
    if (traversal > RTB_TRAVERSAL_SAH16) return fail(RTB_ERR_UNSUPPORTED, "traversal mode %u not supported", traversal);
    if (n == 0) return RTB_OK;
    std::lock_guard<std::mutex> lock(scene->mutex);
    RTB_CUDA(cudaSetDevice(scene->device));
    if (traversal != RTB_TRAVERSAL_REFERENCE) {  // ray queries in reference order walk the (min, max) layout built at creation
        const int rc = ensure_layouts(scene, traversal);
        if (rc != RTB_OK) return rc;
    }
    if (traversal == RTB_TRAVERSAL_SAH16 && !scene->dev.pk_nodes) traversal = RTB_TRAVERSAL_SAH;  // too large to pack
    RtbRay* d_rays = nullptr;
    RtbHit* d_hits = nullptr;
    RTB_CUDA(cudaMalloc(&d_rays, n * sizeof(RtbRay)));
    cudaError_t e = cudaMalloc(&d_hits, n * sizeof(RtbHit));
    if (e == cudaSuccess) e = cudaMemcpy(d_rays, rays, n * sizeof(RtbRay), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_trace(scene->dev, d_rays, n, d_hits, traversal, 0);
    if (e == cudaSuccess) e = cudaMemcpy(hits_out, d_hits, n * sizeof(RtbHit), cudaMemcpyDeviceToHost);
    cudaFree(d_rays);
    cudaFree(d_hits);
    if (e != cudaSuccess) return cuda_fail(e, "rtb_trace_rays");
    return RTB_OK;
}

// ------------------------------------------------------------------ render
static DevCamera lower_camera(const RtbCamera* c) {
    DevCamera d;
    d.center = make_float3(c->center[0], c->center[1], c->center[2]);
    d.pixel00 = make_float3(c->pixel00_loc[0], c->pixel00_loc[1], c->pixel00_loc[2]);
    d.du = make_float3(c->pixel_delta_u[0], c->pixel_delta_u[1], c->pixel_delta_u[2]);
    d.dv = make_float3(c->pixel_delta_v[0], c->pixel_delta_v[1], c->pixel_delta_v[2]);
    d.ddu = make_float3(c->defocus_disk_u[0], c->defocus_disk_u[1], c->defocus_disk_u[2]);
    d.ddv = make_float3(c->defocus_disk_v[0], c->defocus_disk_v[1], c->defocus_disk_v[2]);
    d.background = make_float3(c->background[0], c->background[1], c->background[2]);
    d.defocus_angle = c->defocus_angle;
    d.width = c->image_width;
    d.height = c->image_height;
    d.max_depth = c->max_depth;
    d.background_mode = c->background_mode;
    return d;
}

static int check_render_args(RtbScene* scene, const RtbCamera* cam, const RtbRenderOptions* opt) {
    if (!scene || !cam || !opt) return fail(RTB_ERR_INVALID_ARGUMENT, "scene/camera/options is NULL");
    if (cam->image_width == 0 || cam->image_height == 0) return fail(RTB_ERR_INVALID_ARGUMENT, "empty image");
    const uint64_t size = (uint64_t)cam->image_width * cam->image_height;
    if (size > 0xffffffffull) return fail(RTB_ERR_INVALID_ARGUMENT, "image too large");
    if ((uint64_t)opt->pixel_begin + opt->pixel_count > size) return fail(RTB_ERR_INVALID_ARGUMENT, "pixel range out of bounds");
    if (opt->tile_world > 0 && opt->tile_rank >= opt->tile_world) return fail(RTB_ERR_INVALID_ARGUMENT, "tile_rank >= tile_world");
    if (opt->integrator != RTB_INTEGRATOR_MEGAKERNEL && opt->integrator != RTB_INTEGRATOR_WAVEFRONT)
        return fail(RTB_ERR_INVALID_ARGUMENT, "unknown integrator %u", opt->integrator);
    if (opt->traversal > RTB_TRAVERSAL_SAH16) return fail(RTB_ERR_UNSUPPORTED, "traversal mode %u not supported", opt->traversal);
    if (cam->background_mode > RTB_BACKGROUND_SKY) return fail(RTB_ERR_INVALID_ARGUMENT, "unknown background mode");
    return RTB_OK;
}

// Pixels this call renders: the flat range [pixel_begin, +pixel_count) intersected with the 32x8 tiles t for which
// t % tile_world == tile_rank.
static uint64_t owned_pixels(const RtbCamera* cam, const RtbRenderOptions* opt) {
    const uint64_t W = cam->image_width, H = cam->image_height;
    const uint64_t lo = opt->pixel_begin;
    const uint64_t hi = (opt->pixel_begin == 0 && opt->pixel_count == 0) ? W * H : lo + opt->pixel_count;
    const uint64_t world = opt->tile_world ? opt->tile_world : 1u;
    if (world == 1) return hi - lo;
    const uint64_t tiles_x = (W + kTileW - 1) / kTileW, tiles_y = (H + kTileH - 1) / kTileH;
    uint64_t n = 0;
    for (uint64_t t = opt->tile_rank; t < tiles_x * tiles_y; t += world) {
        const uint64_t x0 = (t % tiles_x) * kTileW, y0 = (t / tiles_x) * kTileH;
        const uint64_t x1 = x0 + kTileW < W ? x0 + kTileW : W, y1 = y0 + kTileH < H ? y0 + kTileH : H;
        for (uint64_t y = y0; y < y1; ++y) {
            const uint64_t a = y * W + x0 > lo ? y * W + x0 : lo, b = y * W + x1 < hi ? y * W + x1 : hi;
            if (b > a) n += b - a;
        }
    }
    return n;
}

// Core: enqueue the kernels for samples [begin, begin+count) on `stream`.  Caller holds the lock.
static int render_enqueue(RtbScene* scene, const RtbCamera* cam, const RtbRenderOptions* opt, uint32_t begin,
                          uint32_t count, float* d_accum, cudaStream_t stream, LaunchInfo* info) {
    // the megakernel in reference order walks the (min, max) layout built at creation; everything else a per-octant one
    if (opt->traversal != RTB_TRAVERSAL_REFERENCE || opt->integrator == RTB_INTEGRATOR_WAVEFRONT) {
        const int rc = ensure_layouts(scene, opt->traversal);
        if (rc != RTB_OK) return rc;
    }
    RenderParams p{};
    p.scene = scene->dev;
    p.cam = lower_camera(cam);
    p.accum = reinterpret_cast<float4*>(d_accum);
    p.seed = make_uint2((uint32_t)opt->seed, (uint32_t)(opt->seed >> 32));
    p.sample_begin = begin;
    p.sample_count = count;
    const uint32_t size = cam->image_width * cam->image_height;
    p.pixel_begin = opt->pixel_begin;
    p.pixel_end = (opt->pixel_begin == 0 && opt->pixel_count == 0) ? size : opt->pixel_begin + opt->pixel_count;
    p.tile_rank = opt->tile_rank;
    p.tile_world = opt->tile_world ? opt->tile_world : 1u;
    p.ordered = opt->traversal;
    if (p.ordered == RTB_TRAVERSAL_SAH16 && !scene->dev.pk_nodes) p.ordered = RTB_TRAVERSAL_SAH;  // too large to pack
    const bool count_work = (opt->flags & RTB_FLAG_COUNT_WORK) != 0;
    p.counters = count_work ? scene->d_counters : nullptr;
    if (opt->integrator == RTB_INTEGRATOR_WAVEFRONT) {
        if (!scene->wavefront) {
            scene->wavefront = wavefront_create();
            if (!scene->wavefront) return fail(RTB_ERR_OUT_OF_MEMORY, "wavefront state allocation failed");
        }
        const cudaError_t e = wavefront_render(scene->wavefront, p, count_work, stream, info);
        if (e != cudaSuccess) return cuda_fail(e, "wavefront_render");
        return RTB_OK;
    }
    const cudaError_t e = launch_megakernel(p, scene->nodes_fit_smem, count_work, stream, info);
    if (e != cudaSuccess) return cuda_fail(e, "launch_megakernel");
    return RTB_OK;
}

static int render_device_locked(RtbScene* scene, const RtbCamera* cam, const RtbRenderOptions* opt, float* d_accum,
                                cudaStream_t stream, RtbRenderStats* stats) {
    const uint32_t total = opt->sample_count ? opt->sample_count : cam->samples_per_pixel;
    const uint32_t per_launch = opt->samples_per_launch ? opt->samples_per_launch : total;
    const bool count_work = (opt->flags & RTB_FLAG_COUNT_WORK) != 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (stats) {
        cudaError_t e = cudaEventCreate(&ev0);
        if (e == cudaSuccess) e = cudaEventCreate(&ev1);
        if (e == cudaSuccess && count_work) e = cudaMemsetAsync(scene->d_counters, 0, 4 * sizeof(unsigned long long), stream);
        if (e == cudaSuccess) e = cudaEventRecord(ev0, stream);
        if (e != cudaSuccess) {
            if (ev0) cudaEventDestroy(ev0);
            if (ev1) cudaEventDestroy(ev1);
            return cuda_fail(e, "render timers");
        }
    }
    LaunchInfo info;
    int rc = RTB_OK;
    for (uint32_t done = 0; done < total && rc == RTB_OK;) {
        const uint32_t n = (total - done < per_launch) ? total - done : per_launch;
        rc = render_enqueue(scene, cam, opt, opt->sample_begin + done, n, d_accum, stream, &info);
        done += n;
    }
    if (stats) {
        if (rc == RTB_OK) {
            cudaError_t e = cudaEventRecord(ev1, stream);
            if (e == cudaSuccess) e = cudaEventSynchronize(ev1);
            float ms = 0;
            if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, ev0, ev1);
            std::memset(stats, 0, sizeof(*stats));
            stats->device_ms = ms;
            stats->n_launches = info.n_launches;
            stats->n_paths = owned_pixels(cam, opt) * total;
            if (e == cudaSuccess && count_work) {
                unsigned long long c[4];
                e = cudaMemcpy(c, scene->d_counters, sizeof(c), cudaMemcpyDeviceToHost);
                stats->n_rays = c[0];
                stats->n_box_tests = c[1];
                stats->n_object_tests = c[2];
                stats->n_hits = c[3];
            }
            if (e != cudaSuccess) rc = cuda_fail(e, "render stats");
        }
        cudaEventDestroy(ev0);
        cudaEventDestroy(ev1);
    }
    return rc;
}

extern "C" int rtb_render_device(RtbScene* scene, const RtbCamera* camera, const RtbRenderOptions* options,
                                 float* d_accum, void* cuda_stream, RtbRenderStats* stats) {
    int rc = check_render_args(scene, camera, options);
    if (rc != RTB_OK) return rc;
    if (!d_accum) return fail(RTB_ERR_INVALID_ARGUMENT, "d_accum is NULL");
    std::lock_guard<std::mutex> lock(scene->mutex);
    RTB_CUDA(cudaSetDevice(scene->device));
    return render_device_locked(scene, camera, options, d_accum, static_cast<cudaStream_t>(cuda_stream), stats);
}

extern "C" int rtb_resolve_device(const float* d_accum, uint8_t* d_rgba, uint64_t n_pixels, float n_samples_override,
                                  int device, void* cuda_stream) {
    if (n_pixels && (!d_accum || !d_rgba)) return fail(RTB_ERR_INVALID_ARGUMENT, "d_accum/d_rgba is NULL");
    RTB_CUDA(cudaSetDevice(device));
    const cudaError_t e = launch_resolve(reinterpret_cast<const float4*>(d_accum), reinterpret_cast<uchar4*>(d_rgba),
                                         n_pixels, n_samples_override, static_cast<cudaStream_t>(cuda_stream));
    if (e != cudaSuccess) return cuda_fail(e, "launch_resolve");
    return RTB_OK;
}

// ---- multi-GPU exchange over peer memory ------------------------------------------------------------------------
extern "C" int rtb_buffer_alloc(int device, uint64_t bytes, void** device_ptr_out) {
    if (!device_ptr_out) return fail(RTB_ERR_INVALID_ARGUMENT, "device_ptr_out is NULL");
    *device_ptr_out = nullptr;
    RTB_CUDA(cudaSetDevice(device));
    void* p = nullptr;
    const cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(exchange buffer)");
    *device_ptr_out = p;
    return RTB_OK;
}

extern "C" int rtb_buffer_free(int device, void* device_ptr) {
    if (!device_ptr) return RTB_OK;
    RTB_CUDA(cudaSetDevice(device));
    RTB_CUDA(cudaFree(device_ptr));
    return RTB_OK;
}

extern "C" int rtb_ipc_export(int device, const void* device_ptr, RtbIpcHandle* handle_out) {
    if (!device_ptr || !handle_out) return fail(RTB_ERR_INVALID_ARGUMENT, "device_ptr/handle_out is NULL");
    static_assert(sizeof(cudaIpcMemHandle_t) == sizeof(RtbIpcHandle), "RtbIpcHandle must hold a cudaIpcMemHandle_t");
    RTB_CUDA(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    RTB_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(device_ptr)));
    std::memcpy(handle_out->bytes, &h, sizeof(h));
    return RTB_OK;
}

extern "C" int rtb_ipc_open(int device, const RtbIpcHandle* handle, void** device_ptr_out) {
    if (!handle || !device_ptr_out) return fail(RTB_ERR_INVALID_ARGUMENT, "handle/device_ptr_out is NULL");
    *device_ptr_out = nullptr;
    RTB_CUDA(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle->bytes, sizeof(h));
    void* p = nullptr;
    RTB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *device_ptr_out = p;
    return RTB_OK;
}

extern "C" int rtb_ipc_close(int device, void* device_ptr) {
    if (!device_ptr) return RTB_OK;
    RTB_CUDA(cudaSetDevice(device));
    RTB_CUDA(cudaIpcCloseMemHandle(device_ptr));
    return RTB_OK;
}

extern "C" int rtb_exchange_slice(uint64_t n_pixels, uint32_t world, uint32_t rank, uint32_t root, uint64_t* begin_out,
                                  uint64_t* end_out) {
    if (!begin_out || !end_out) return fail(RTB_ERR_INVALID_ARGUMENT, "begin_out/end_out is NULL");
    if (world == 0 || rank >= world || root >= world)
        return fail(RTB_ERR_INVALID_ARGUMENT, "rank %u / root %u out of range (world %u)", rank, root, world);
    // Which ranks combine (see rtb.h for the policy); k = number of combining ranks before this one.
    auto works = [&](uint32_t r) { return world == 1 ? true : (world == 2 ? r == root : r != root); };
    uint32_t workers = 0, k = 0;
    for (uint32_t r = 0; r < world; ++r) {
        if (!works(r)) continue;
        ++workers;
        if (r < rank) ++k;
    }
    uint64_t per = (n_pixels + workers - 1) / workers;
    per = (per + 255u) / 256u * 256u;
    uint64_t b = per * k, e = per * (k + (works(rank) ? 1ull : 0ull));
    if (b > n_pixels) b = n_pixels;
    if (e > n_pixels) e = n_pixels;
    *begin_out = b;
    *end_out = e;
    return RTB_OK;
}

extern "C" int rtb_exchange_resolve(const float* const* peer_accum, uint32_t world, uint32_t rank, uint32_t root,
                                    float* root_accum_out, uint8_t* root_rgba_out, uint64_t n_pixels,
                                    float samples_per_pixel, int device, void* cuda_stream) {
    if (world == 0 || world > kMaxPeers || rank >= world || root >= world)
        return fail(RTB_ERR_INVALID_ARGUMENT, "world %u / rank %u / root %u out of range (max %u ranks)", world, rank, root,
                    kMaxPeers);
    if (!peer_accum || !root_accum_out || !root_rgba_out) return fail(RTB_ERR_INVALID_ARGUMENT, "NULL buffer");
    if (!(samples_per_pixel > 0.0f)) return fail(RTB_ERR_INVALID_ARGUMENT, "samples_per_pixel must be > 0");
    PeerAccums peers{};
    for (uint32_t r = 0; r < world; ++r) {
        if (!peer_accum[r]) return fail(RTB_ERR_INVALID_ARGUMENT, "peer_accum[%u] is NULL", r);
        peers.p[r] = reinterpret_cast<const float4*>(peer_accum[r]);
    }
    uint64_t begin = 0, end = 0;
    const int rc = rtb_exchange_slice(n_pixels, world, rank, root, &begin, &end);
    if (rc != RTB_OK) return rc;
    RTB_CUDA(cudaSetDevice(device));
    const cudaError_t e = launch_exchange_resolve(peers, world, reinterpret_cast<float4*>(root_accum_out),
                                                  reinterpret_cast<uchar4*>(root_rgba_out), begin, end, samples_per_pixel,
                                                  static_cast<cudaStream_t>(cuda_stream));
    if (e != cudaSuccess) return cuda_fail(e, "launch_exchange_resolve");
    return RTB_OK;
}

extern "C" int rtb_resolve(const float* accum, uint8_t* rgba, uint64_t n_pixels, float n_samples_override, int device) {
    if (n_pixels && (!accum || !rgba)) return fail(RTB_ERR_INVALID_ARGUMENT, "accum/rgba is NULL");
    if (n_pixels == 0) return RTB_OK;
    int ndev = 0;
    int rc = rtb_device_count(&ndev);
    if (rc != RTB_OK) return rc;
    RTB_CUDA(cudaSetDevice(device));
    float* d_acc = nullptr;
    uint8_t* d_rgba = nullptr;
    RTB_CUDA(cudaMalloc(&d_acc, n_pixels * 16));
    cudaError_t e = cudaMalloc(&d_rgba, n_pixels * 4);
    if (e == cudaSuccess) e = cudaMemcpy(d_acc, accum, n_pixels * 16, cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
        e = launch_resolve(reinterpret_cast<const float4*>(d_acc), reinterpret_cast<uchar4*>(d_rgba), n_pixels,
                           n_samples_override, 0);
    if (e == cudaSuccess) e = cudaMemcpy(rgba, d_rgba, n_pixels * 4, cudaMemcpyDeviceToHost);
    cudaFree(d_acc);
    cudaFree(d_rgba);
    if (e != cudaSuccess) return cuda_fail(e, "rtb_resolve");
    return RTB_OK;
}

// Host-buffer render: upload accum, render, resolve, download.  `progress` (optional) is called
// after every batch with the number of samples done; returning non-zero cancels.
template <class Progress>
static int render_host(RtbScene* scene, const RtbCamera* cam, const RtbRenderOptions* opt, float* accum,
                       uint8_t* rgba, RtbRenderStats* stats, bool refresh_each_batch, Progress progress) {
    int rc = check_render_args(scene, cam, opt);
    if (rc != RTB_OK) return rc;
    if (!accum) return fail(RTB_ERR_INVALID_ARGUMENT, "accum is NULL");
    std::lock_guard<std::mutex> lock(scene->mutex);
    RTB_CUDA(cudaSetDevice(scene->device));
    const uint64_t npx = (uint64_t)cam->image_width * cam->image_height;
    if (scene->stage_pixels < npx) {  // (re)size the staging buffers; they live as long as the scene
        cudaFree(scene->stage_accum);
        cudaFree(scene->stage_rgba);
        scene->stage_accum = nullptr;
        scene->stage_rgba = nullptr;
        scene->stage_pixels = 0;
        RTB_CUDA(cudaMalloc(&scene->stage_accum, npx * 16));
        RTB_CUDA(cudaMalloc(&scene->stage_rgba, npx * 4));
        scene->stage_pixels = npx;
    }
    float* d_acc = scene->stage_accum;
    uint8_t* d_rgba = rgba ? scene->stage_rgba : nullptr;
    cudaError_t e = cudaMemcpy(d_acc, accum, npx * 16, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return cuda_fail(e, "rtb_render upload");
    const uint32_t total = opt->sample_count ? opt->sample_count : cam->samples_per_pixel;
    const uint32_t per_batch = opt->samples_per_launch ? opt->samples_per_launch : total;
    RtbRenderStats acc_stats{};
    bool cancelled = false;
    for (uint32_t done = 0; done < total && rc == RTB_OK && !cancelled;) {
        RtbRenderOptions o = *opt;
        o.sample_begin = opt->sample_begin + done;
        o.sample_count = (total - done < per_batch) ? total - done : per_batch;
        o.samples_per_launch = o.sample_count;
        if (o.sample_count == 0) break;
        RtbRenderStats st{};
        rc = render_device_locked(scene, cam, &o, d_acc, 0, &st);
        if (rc != RTB_OK) break;
        done += o.sample_count;
        acc_stats.n_paths += st.n_paths;
        acc_stats.n_rays += st.n_rays;
        acc_stats.n_box_tests += st.n_box_tests;
        acc_stats.n_object_tests += st.n_object_tests;
        acc_stats.n_hits += st.n_hits;
        acc_stats.device_ms += st.device_ms;
        acc_stats.n_launches += st.n_launches;
        const bool last = done >= total;
        if (refresh_each_batch || last) {
            if (d_rgba) {
                e = launch_resolve(reinterpret_cast<const float4*>(d_acc), reinterpret_cast<uchar4*>(d_rgba), npx, 0.0f, 0);
                acc_stats.n_launches += 1;
                if (e == cudaSuccess) e = cudaMemcpy(rgba, d_rgba, npx * 4, cudaMemcpyDeviceToHost);
            }
            if (e == cudaSuccess) e = cudaMemcpy(accum, d_acc, npx * 16, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) {
                rc = cuda_fail(e, "rtb_render download");
                break;
            }
        }
        if (progress(done) && !last) {
            cancelled = true;
            if (!refresh_each_batch) {  // hand back what has been rendered so far
                if (d_rgba) {
                    e = launch_resolve(reinterpret_cast<const float4*>(d_acc), reinterpret_cast<uchar4*>(d_rgba), npx, 0.0f, 0);
                    if (e == cudaSuccess) e = cudaMemcpy(rgba, d_rgba, npx * 4, cudaMemcpyDeviceToHost);
                }
                if (e == cudaSuccess) e = cudaMemcpy(accum, d_acc, npx * 16, cudaMemcpyDeviceToHost);
                if (e != cudaSuccess) rc = cuda_fail(e, "rtb_render download");
            }
        }
    }
    if (stats) *stats = acc_stats;
    if (rc == RTB_OK && cancelled) return fail(RTB_ERR_CANCELLED, "render cancelled");
    return rc;
}

extern "C" int rtb_render(RtbScene* scene, const RtbCamera* camera, const RtbRenderOptions* options, float* accum,
                          uint8_t* rgba, RtbRenderStats* stats) {
    return render_host(scene, camera, options, accum, rgba, stats, false, [](uint32_t) { return 0; });
}

// ------------------------------------------------------------------ async jobs
extern "C" int rtb_render_async(RtbScene* scene, const RtbCamera* camera, const RtbRenderOptions* options,
                                float* accum, uint8_t* rgba, RtbJob** job_out) {
    if (!job_out) return fail(RTB_ERR_INVALID_ARGUMENT, "job_out is NULL");
    *job_out = nullptr;
    int rc = check_render_args(scene, camera, options);
    if (rc != RTB_OK) return rc;
    if (!accum) return fail(RTB_ERR_INVALID_ARGUMENT, "accum is NULL");
    RtbJob* job = new (std::nothrow) RtbJob();
    if (!job) return fail(RTB_ERR_OUT_OF_MEMORY, "host allocation failed");
    const RtbCamera cam = *camera;
    RtbRenderOptions opt = *options;
    job->samples_total = opt.sample_count ? opt.sample_count : cam.samples_per_pixel;
    if (opt.samples_per_launch == 0) opt.samples_per_launch = 1;  // the reference refreshes after every sample
    job->worker = std::thread([=]() {
        RtbRenderStats st{};
        const int r = render_host(scene, &cam, &opt, accum, rgba, &st, true, [job](uint32_t done) {
            job->samples_done.store(done, std::memory_order_release);
            return job->cancel.load(std::memory_order_acquire);
        });
        job->status = r;
        if (r != RTB_OK) job->error = g_last_error;
        job->stats = st;
        job->running.store(0, std::memory_order_release);
    });
    *job_out = job;
    return RTB_OK;
}

extern "C" int rtb_job_progress(RtbJob* job, uint32_t* samples_done, uint32_t* samples_total, int* running) {
    if (!job) return fail(RTB_ERR_INVALID_ARGUMENT, "job is NULL");
    if (samples_done) *samples_done = job->samples_done.load(std::memory_order_acquire);
    if (samples_total) *samples_total = job->samples_total;
    if (running) *running = job->running.load(std::memory_order_acquire);
    return RTB_OK;
}

extern "C" int rtb_job_cancel(RtbJob* job) {
    if (!job) return fail(RTB_ERR_INVALID_ARGUMENT, "job is NULL");
    job->cancel.store(1, std::memory_order_release);
    return RTB_OK;
}

extern "C" int rtb_job_wait(RtbJob* job, RtbRenderStats* stats) {
    if (!job) return fail(RTB_ERR_INVALID_ARGUMENT, "job is NULL");
    if (job->worker.joinable()) job->worker.join();
    if (stats) *stats = job->stats;
    if (job->status != RTB_OK) g_last_error = job->error;
    return job->status;
}

extern "C" int rtb_job_destroy(RtbJob* job) {
    if (!job) return RTB_OK;
    job->cancel.store(1, std::memory_order_release);
    if (job->worker.joinable()) job->worker.join();
    delete job;
    return RTB_OK;
}

// ------------------------------------------------------------------ RNG self-test
extern "C" int rtb_philox_device_selftest(const uint32_t* counters4, const uint32_t* key2, uint32_t n, uint32_t* out4,
                                          int device) {
    if (n && (!counters4 || !key2 || !out4)) return fail(RTB_ERR_INVALID_ARGUMENT, "NULL argument");
    if (n == 0) return RTB_OK;
    int ndev = 0;
    int rc = rtb_device_count(&ndev);
    if (rc != RTB_OK) return rc;
    RTB_CUDA(cudaSetDevice(device));
    uint4 *d_in = nullptr, *d_out = nullptr;
    RTB_CUDA(cudaMalloc(&d_in, (size_t)n * 16));
    cudaError_t e = cudaMalloc(&d_out, (size_t)n * 16);
    if (e == cudaSuccess) e = cudaMemcpy(d_in, counters4, (size_t)n * 16, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_philox_selftest(d_in, make_uint2(key2[0], key2[1]), n, d_out, 0);
    if (e == cudaSuccess) e = cudaMemcpy(out4, d_out, (size_t)n * 16, cudaMemcpyDeviceToHost);
    cudaFree(d_in);
    cudaFree(d_out);
    if (e != cudaSuccess) return cuda_fail(e, "rtb_philox_device_selftest");
    return RTB_OK;
}

// ------------------------------------------------------------------ FP32 peak (roofline denominator)
extern "C" int rtb_measure_fp32_peak(int device, double* tflops_out) {
    if (!tflops_out) return fail(RTB_ERR_INVALID_ARGUMENT, "tflops_out is NULL");
    int ndev = 0;
    int rc = rtb_device_count(&ndev);
    if (rc != RTB_OK) return rc;
    RTB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    RTB_CUDA(cudaGetDeviceProperties(&prop, device));
    const uint32_t grid = (uint32_t)prop.multiProcessorCount * 8u;
    const uint32_t iters = 1u << 15;
    float* d_out = nullptr;
    RTB_CUDA(cudaMalloc(&d_out, (size_t)grid * 256 * sizeof(float)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    cudaError_t e = cudaSuccess;
    for (int rep = 0; rep < 6 && e == cudaSuccess; ++rep) {
        cudaEventRecord(e0, 0);
        e = launch_ffma_peak(d_out, grid, iters, 0);
        cudaEventRecord(e1, 0);
        if (e == cudaSuccess) e = cudaEventSynchronize(e1);
        float ms = 0;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        const double tf = (double)grid * 256.0 * iters * 8.0 * 2.0 / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;  // rep 0 is the warm-up
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    if (e != cudaSuccess) return cuda_fail(e, "rtb_measure_fp32_peak");
    *tflops_out = best;
    return RTB_OK;
}
