// Host side of libb200pt: scene packing (reference layout -> B200 device layout) and the native
// binned-SAH BVH builder.  Plain C++17, no CUDA calls.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../include/b200pt.h"

namespace {

thread_local std::string g_host_error_storage;
#define g_host_error (g_host_error_storage)

constexpr int kRefNone = (int)0x80000000;
// Leaf reference: ~(prim_idx << 5 | shading_class << 2 | prim_type).  The shading class rides along so
// that the kernels can tell a medium boundary and pick the material bin of a hit without touching the
// material tables: 1 emissive, 2 Lambert with a solid colour, 3 Lambert with a checker / image / noise
// texture, 4 metal, 5 dielectric, 6 constant-medium boundary, 7 everything else (isotropic, subsurface);
// 0 is reserved for "miss" (pt_device.cuh PT_CLS_*).
constexpr int kMaxPrimsPerType = 1 << 25;   // bit 30 of a leaf reference stays set: the traversal clears it to flag an exact-t tie (kRefTie)
inline int leaf_ref(int prim_type, int prim_idx, int cls) { return ~((prim_idx << 5) | ((cls & 7) << 2) | prim_type); }
inline int shading_class(const PtHostMaterials& m, int i) {
    if (m.is_medium[i] > 0) return 6;
    switch (m.material_type[i]) {
        case 0: return m.texture_type[i] == 0 ? 2 : 3;
        case 1: return 4;
        case 2: return 5;
        case 3: return 1;
        default: return 7;
    }
}
inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

struct MatRec {
    uint32_t w[12];
    bool operator==(const MatRec& o) const { return memcmp(w, o.w, sizeof w) == 0; }
};
struct MatHash {
    size_t operator()(const MatRec& m) const {
        uint64_t h = 1469598103934665603ull;
        for (int i = 0; i < 12; i++) { h ^= m.w[i]; h *= 1099511628211ull; }
        return (size_t)h;
    }
};
inline uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

// One 48-byte record per *distinct* material.  Only the fields the kernels read for that material
// type are kept (the rest is zeroed so that equal materials compare equal):
//   w0 header = mat_type | tex_type<<4 | is_medium<<8 | (image_idx+1)<<16
//   w1 scalar = Lambert/isotropic: texture_scale; metal: fuzz; dielectric: ir
//   w2..4 c1  = Lambert/isotropic: texture_color1; metal: albedo; emissive: emit colour
//   w5..7 c2  = checker: texture_color2
//   w8 density, w9..11 medium albedo (medium boundaries only)
MatRec make_mat(const PtHostMaterials& m, int i) {
    MatRec r;
    memset(&r, 0, sizeof r);
    const int mt = m.material_type[i];
    int tt = 0, img = -1;
    float scalar = 0.0f;
    const float* c1 = nullptr;
    const float* c2 = nullptr;
    if (mt == 0 || mt == 4) {
        tt = m.texture_type[i];
        c1 = m.color1 + 3 * i;
        if (tt == 1) { c2 = m.color2 + 3 * i; scalar = m.texture_scale[i]; }
        else if (tt == 2) { img = m.image_idx[i]; c1 = nullptr; }
        else if (tt == 3) { scalar = m.texture_scale[i]; }
    } else if (mt == 1) {
        scalar = m.fuzz[i];
        c1 = m.albedo + 3 * i;
    } else if (mt == 2) {
        scalar = m.ir[i];
    } else if (mt == 3) {
        c1 = m.emit + 3 * i;
    } else if (mt == 5) {                  // subsurface_simple (full feature mode): scalar = scatter distance
        scalar = m.fuzz[i];
        c1 = m.albedo + 3 * i;
    } else if (mt == 6) {                  // subsurface_volumetric: scalar = sigma_s, c2 = (sigma_a, g, 0)
        scalar = m.fuzz[i];
        c1 = m.albedo + 3 * i;
    }
    const int is_med = m.is_medium[i] > 0 ? 1 : 0;
    r.w[0] = (uint32_t)(mt & 15) | ((uint32_t)(tt & 15) << 4) | ((uint32_t)is_med << 8) | ((uint32_t)(img + 1) << 16);
    r.w[1] = fbits(scalar);
    if (c1) for (int k = 0; k < 3; k++) r.w[2 + k] = fbits(c1[k]);
    if (c2) for (int k = 0; k < 3; k++) r.w[5 + k] = fbits(c2[k]);
    if (mt == 6) { r.w[5] = fbits(m.ir[i]); r.w[6] = fbits(m.texture_scale[i]); }
    if (is_med) {
        r.w[8] = fbits(m.medium_density[i]);
        for (int k = 0; k < 3; k++) r.w[9 + k] = fbits(m.medium_albedo[3 * i + k]);
    }
    return r;
}

struct MatTable {
    std::vector<MatRec> recs;
    std::vector<uint32_t> ids;   // [spheres | triangles | quads]
    int has_media = 0;
};

void build_materials(const PtHostScene* s, MatTable& out) {
    std::unordered_map<MatRec, uint32_t, MatHash> seen;
    const int counts[3] = {s->num_spheres, s->num_triangles, s->num_quads};
    for (int pt = 0; pt < 3; pt++) {
        for (int i = 0; i < counts[pt]; i++) {
            MatRec r = make_mat(s->mat[pt], i);
            if (r.w[0] & 0x100u) out.has_media = 1;
            auto it = seen.find(r);
            uint32_t id;
            if (it == seen.end()) {
                id = (uint32_t)out.recs.size();
                seen.emplace(r, id);
                out.recs.push_back(r);
            } else {
                id = it->second;
            }
            out.ids.push_back(id);
        }
    }
}

// Inner nodes of the reference tree in breadth-first order (so that "the first K nodes" is the
// top of the tree, which the kernels stage in shared memory).
struct Collapsed {
    std::vector<int> order;     // new index -> reference node index
    std::vector<int> new_index; // reference node index -> new index (or -1)
    int depth = 0;
    bool root_is_leaf = false;
};

int collapse(const PtHostScene* s, Collapsed& c) {
    const int n = s->num_nodes;
    c.new_index.assign(n > 0 ? n : 0, -1);
    if (n == 0) return PT_OK;
    if (s->bvh_prim_idx[0] >= 0) { c.root_is_leaf = true; c.depth = 1; return PT_OK; }
    std::vector<int> level{0}, next;
    while (!level.empty()) {
        c.depth++;
        next.clear();
        for (int ni : level) {
            c.new_index[ni] = (int)c.order.size();
            c.order.push_back(ni);
            const int ch[2] = {s->bvh_left[ni], s->bvh_right[ni]};
            if (ch[0] < 0 && ch[1] < 0) { g_host_error = "BVH inner node without children"; return PT_ERR_INVALID_ARG; }
            for (int k = 0; k < 2; k++) {
                if (ch[k] < 0) continue;
                if (ch[k] >= n) { g_host_error = "BVH child index out of range"; return PT_ERR_INVALID_ARG; }
                if (s->bvh_prim_idx[ch[k]] < 0) next.push_back(ch[k]);
            }
        }
        level.swap(next);
        if (c.depth > 4096) { g_host_error = "BVH has a cycle"; return PT_ERR_INVALID_ARG; }
    }
    c.depth++; // leaf level
    return PT_OK;
}

// Everything the packer dereferences, checked once up front: a caller-built scene (the public
// `compiled_scene=` path, or any C caller) must come back as PT_ERR_INVALID_ARG, not as a crash.
int validate_scene(const PtHostScene* s) {
    auto fail = [](const char* msg) { g_host_error = std::string("scene: ") + msg; return (int)PT_ERR_INVALID_ARG; };
    if (s->num_spheres < 0 || s->num_quads < 0 || s->num_triangles < 0 || s->num_nodes < 0 || s->num_images < 0)
        return fail("negative count");
    if (s->num_spheres > 0 && !s->sphere_data) return fail("sphere_data is NULL");
    if (s->num_quads > 0 && (!s->quad_Q || !s->quad_u || !s->quad_v || !s->quad_normal || !s->quad_D || !s->quad_w))
        return fail("a quad array is NULL");
    if (s->num_triangles > 0 && (!s->tri_v0 || !s->tri_edge1 || !s->tri_edge2 || !s->tri_normal))
        return fail("a triangle array is NULL");
    if (s->full_features && s->num_spheres > 0 && !s->sphere_motion) return fail("full feature mode without sphere_motion");
    for (int i = 0; i < s->num_images; i++)
        if (!s->images[i] || s->image_w[i] <= 0 || s->image_h[i] <= 0) return fail("image texture without pixels");
    const int counts[3] = {s->num_spheres, s->num_triangles, s->num_quads};
    for (int pt = 0; pt < 3; pt++) {
        const PtHostMaterials& m = s->mat[pt];
        if (counts[pt] == 0) continue;
        if (!m.material_type || !m.albedo || !m.fuzz || !m.ir || !m.emit || !m.texture_type || !m.texture_scale ||
            !m.color1 || !m.color2 || !m.image_idx || !m.is_medium || !m.medium_density || !m.medium_albedo)
            return fail("a material array is NULL (every per-primitive material array is required)");
        for (int i = 0; i < counts[pt]; i++) {
            if (m.material_type[i] < 0 || m.material_type[i] > 6) return fail("material_type outside 0..6");
            if (m.texture_type[i] < 0 || m.texture_type[i] > 3) return fail("texture_type outside 0..3");
            if ((m.material_type[i] == 0 || m.material_type[i] == 4) && m.texture_type[i] == 2 &&
                (m.image_idx[i] < -1 || m.image_idx[i] >= s->num_images))
                return fail("texture_image_idx outside the image table");
        }
    }
    if (s->num_nodes > 0) {
        if (!s->bvh_bbox_min || !s->bvh_bbox_max || !s->bvh_left || !s->bvh_right || !s->bvh_prim_type || !s->bvh_prim_idx)
            return fail("a BVH array is NULL");
        for (int i = 0; i < s->num_nodes; i++) {
            if (s->bvh_prim_idx[i] < 0) continue;
            const int pt = s->bvh_prim_type[i];
            if (pt < 0 || pt > 2) return fail("bvh_prim_type outside 0..2");
            if (s->bvh_prim_idx[i] >= counts[pt]) return fail("bvh_prim_idx beyond the primitive count");
        }
    }
    return PT_OK;
}

// ---- light table for next-event estimation: every emissive primitive that is not a medium boundary, with the
// area its sampler needs (quad |u x v|, sphere 4 pi r^2, triangle |e1 x e2| / 2: double arithmetic on the float32
// inputs, narrowed once -- oracle.cpp computes them the same way).  Order: spheres, triangles, quads, by index.
struct LightRec {
    int ref;
    float area, emit[3];
};
void build_lights(const PtHostScene* s, std::vector<LightRec>& out) {
    const int counts[3] = {s->num_spheres, s->num_triangles, s->num_quads};
    for (int pt = 0; pt < 3; pt++) {
        const PtHostMaterials& m = s->mat[pt];
        for (int i = 0; i < counts[pt]; i++) {
            if (m.material_type[i] != 3 || m.is_medium[i] > 0) continue;
            const float* e = m.emit + 3 * i;
            if (!(e[0] > 0.0f || e[1] > 0.0f || e[2] > 0.0f)) continue;
            double area = 0.0;
            if (pt == PT_PRIM_SPHERE) {
                const double r = s->sphere_data[4 * i + 3];
                area = 4.0 * 3.14159265358979323846 * r * r;
            } else {
                const float* a = pt == PT_PRIM_QUAD ? s->quad_u + 3 * i : s->tri_edge1 + 3 * i;
                const float* b = pt == PT_PRIM_QUAD ? s->quad_v + 3 * i : s->tri_edge2 + 3 * i;
                const double cx = (double)a[1] * b[2] - (double)a[2] * b[1], cy = (double)a[2] * b[0] - (double)a[0] * b[2],
                             cz = (double)a[0] * b[1] - (double)a[1] * b[0];
                area = std::sqrt(cx * cx + cy * cy + cz * cz) * (pt == PT_PRIM_QUAD ? 1.0 : 0.5);
            }
            if (!(area > 0.0)) continue;
            out.push_back(LightRec{leaf_ref(pt, i, shading_class(m, i)), (float)area, {e[0], e[1], e[2]}});
        }
    }
}

#ifndef PT_WIDE
#define PT_WIDE 0
#endif
#if PT_WIDE
// ---- 4-wide collapse of the reference's binary tree (SURVEY.md section 8f N1 "optionally wide / collapsed 4-ary").
// A wide node takes a binary inner node's two children and, while it has fewer than four, replaces the inner
// child with the largest surface area by that child's own two children.  Child boxes are the reference's node
// boxes, unchanged; skipped intermediate boxes enclose the ones kept, and the slab test is monotone in the box
// (rounding is monotone), so a ray reaches exactly the leaves whose own boxes it passes -- the same set as in
// the binary tree: closest-hit t is unchanged, primitive ids up to exact-t ties.
struct Wide {
    int kid[4];   // reference node index of each child, -1 = empty slot
    int n;
};
struct Collapsed4 {
    std::vector<Wide> nodes;         // breadth-first
    std::vector<int> wide_of;        // reference inner node -> wide index (for the ones that head a wide node)
    int depth = 0;                   // levels of wide nodes
};
inline float box_area(const PtHostScene* s, int ni) {
    const float dx = s->bvh_bbox_max[3 * ni] - s->bvh_bbox_min[3 * ni], dy = s->bvh_bbox_max[3 * ni + 1] - s->bvh_bbox_min[3 * ni + 1],
                dz = s->bvh_bbox_max[3 * ni + 2] - s->bvh_bbox_min[3 * ni + 2];
    return dx * dy + dy * dz + dz * dx;
}
int collapse4(const PtHostScene* s, Collapsed4& c) {
    const int n = s->num_nodes;
    c.wide_of.assign(n > 0 ? n : 0, -1);
    if (n == 0) return PT_OK;
    auto is_leaf = [&](int ni) { return s->bvh_prim_idx[ni] >= 0; };
    std::vector<int> level{0}, next;
    if (is_leaf(0)) {                       // a one-primitive scene: one wide node with that leaf
        c.nodes.push_back(Wide{{0, -1, -1, -1}, 1});
        c.depth = 1;
        return PT_OK;
    }
    c.wide_of[0] = 0;
    c.nodes.push_back(Wide{{-1, -1, -1, -1}, 0});
    while (!level.empty()) {
        c.depth++;
        if (c.depth > 4096) { g_host_error = "BVH has a cycle"; return PT_ERR_INVALID_ARG; }
        next.clear();
        for (int head : level) {
            std::vector<int> kids;
            for (int ch : {s->bvh_left[head], s->bvh_right[head]}) {
                if (ch < 0) continue;
                if (ch >= n) { g_host_error = "BVH child index out of range"; return PT_ERR_INVALID_ARG; }
                kids.push_back(ch);
            }
            if (kids.empty()) { g_host_error = "BVH inner node without children"; return PT_ERR_INVALID_ARG; }
            while (kids.size() < 4) {
                int best = -1;
                float best_area = -1.0f;
                for (size_t k = 0; k < kids.size(); k++)
                    if (!is_leaf(kids[k]) && box_area(s, kids[k]) > best_area) { best = (int)k; best_area = box_area(s, kids[k]); }
                if (best < 0) break;
                const int open = kids[best], l = s->bvh_left[open], r = s->bvh_right[open];
                if ((l < 0 && r < 0) || l >= n || r >= n) { g_host_error = "BVH inner node with bad children"; return PT_ERR_INVALID_ARG; }
                kids.erase(kids.begin() + best);
                if (l >= 0) kids.push_back(l);
                if (r >= 0) kids.push_back(r);
            }
            Wide w{{-1, -1, -1, -1}, (int)kids.size()};
            for (size_t k = 0; k < kids.size(); k++) {
                w.kid[k] = kids[k];
                if (!is_leaf(kids[k])) {
                    c.wide_of[kids[k]] = (int)c.nodes.size();
                    c.nodes.push_back(Wide{{-1, -1, -1, -1}, 0});
                    next.push_back(kids[k]);
                }
            }
            c.nodes[c.wide_of[head]] = w;
        }
        level.swap(next);
    }
    return PT_OK;
}
#endif

}  // namespace

extern "C" {

// the calling thread's error slot, shared with b200pt.cu (pt_last_error reads it)
std::string* pt_error_slot(void) { return &g_host_error_storage; }

int pt_pack_scene_info(const PtHostScene* s, PtPackedInfo* info) {
    if (!s || !info) { g_host_error = "pt_pack_scene_info: NULL argument"; return PT_ERR_INVALID_ARG; }
    if (s->num_images > PT_MAX_IMAGES) { g_host_error = "too many image textures"; return PT_ERR_CAPACITY; }
    if (s->num_spheres >= kMaxPrimsPerType || s->num_quads >= kMaxPrimsPerType || s->num_triangles >= kMaxPrimsPerType) {
        g_host_error = "more than 2^25 primitives of one type";
        return PT_ERR_CAPACITY;
    }
    memset(info, 0, sizeof *info);
    int rc = validate_scene(s);
    if (rc) return rc;
#if PT_WIDE
    Collapsed4 c;
    rc = collapse4(s, c);
    if (rc) return rc;
    MatTable mt;
    build_materials(s, mt);
    info->num_inner = (int)c.nodes.size();
    info->node_bytes = 128;
#else
    Collapsed c;
    rc = collapse(s, c);
    if (rc) return rc;
    MatTable mt;
    build_materials(s, mt);
    info->num_inner = c.root_is_leaf ? 1 : (int)c.order.size();
    info->node_bytes = 64;
#endif
    info->num_spheres = s->num_spheres;
    info->num_quads = s->num_quads;
    info->num_triangles = s->num_triangles;
    info->num_materials = (int)mt.recs.size();
    info->num_images = s->num_images;
#if PT_WIDE
    info->max_depth_tree = c.depth > 0 ? 3 * c.depth + 1 : 0;     // stack slots a 4-wide traversal can need (3 deferred children per level)
#else
    info->max_depth_tree = c.depth;
#endif
    info->has_media = mt.has_media;
    uint64_t off = 0;
    auto section = [&](uint64_t bytes) { uint64_t o = off; off = align_up(off + bytes, 256); return o; };
    info->off_nodes = section((uint64_t)info->num_inner * (uint64_t)info->node_bytes);
    info->off_spheres = section((uint64_t)s->num_spheres * 16);
    info->off_quads = section((uint64_t)s->num_quads * 64);
    info->off_tris = section((uint64_t)s->num_triangles * 48);
    info->off_prim_mat = section((uint64_t)mt.ids.size() * 4);
    info->off_materials = section((uint64_t)mt.recs.size() * 48);
    info->off_perlin = section(256 * 16 + 768);
    info->full_features = s->full_features & 3;
    info->off_sphere_motion = info->full_features ? section((uint64_t)s->num_spheres * 16) : 0;
    {
        std::vector<LightRec> lights;
        if (info->full_features & 2) build_lights(s, lights);
        info->num_lights = (int)lights.size();
        info->off_lights = section((uint64_t)lights.size() * 32);
        double sum = 0.0;
        for (const LightRec& l : lights) sum += (double)l.area;
        info->nee_r2 = lights.empty() ? 0.0f : (float)(0.25 * (sum / (double)lights.size()));
    }
    info->off_images = off;
    uint64_t ioff = 0;
    for (int i = 0; i < s->num_images; i++) {
        info->image_off[i] = ioff;
        info->image_w[i] = s->image_w[i];
        info->image_h[i] = s->image_h[i];
        ioff = align_up(ioff + (uint64_t)s->image_w[i] * s->image_h[i] * 4, 256);
    }
    off += ioff;
    info->total_bytes = align_up(off, 256);
    if (info->total_bytes == 0) info->total_bytes = 256;
    return PT_OK;
}

int pt_pack_scene(const PtHostScene* s, const PtPackedInfo* info, void* host_blob) {
    if (!s || !info || !host_blob) { g_host_error = "pt_pack_scene: NULL argument"; return PT_ERR_INVALID_ARG; }
    char* base = static_cast<char*>(host_blob);
    int rc = validate_scene(s);
    if (rc) return rc;
    memset(base, 0, info->total_bytes);
#if PT_WIDE
    // ---- nodes, 128 B: (min.x of the 4 children)(max.x)(min.y)(max.y)(min.z)(max.z)(4 child refs)(unused).
    //      An empty slot holds a box no ray can pass (a point at 1e30) and the "none" reference.
    {
        Collapsed4 c;
        rc = collapse4(s, c);
        if (rc) return rc;
        if ((int)c.nodes.size() != info->num_inner) { g_host_error = "pt_pack_scene: info does not match scene"; return PT_ERR_INVALID_ARG; }
        float* nodes = reinterpret_cast<float*>(base + info->off_nodes);
        for (size_t w = 0; w < c.nodes.size(); w++) {
            float* p = nodes + 32 * w;
            for (int k = 0; k < 4; k++) {
                const int ni = c.nodes[w].kid[k];
                int ref = kRefNone;
                for (int a = 0; a < 3; a++) {
                    p[8 * a + k] = ni < 0 ? 1e30f : s->bvh_bbox_min[3 * ni + a];
                    p[8 * a + 4 + k] = ni < 0 ? 1e30f : s->bvh_bbox_max[3 * ni + a];
                }
                if (ni >= 0) {
                    if (s->bvh_prim_idx[ni] >= 0) {
                        const int pt = s->bvh_prim_type[ni], pi = s->bvh_prim_idx[ni];
                        ref = leaf_ref(pt, pi, shading_class(s->mat[pt], pi));
                    } else {
                        ref = c.wide_of[ni];
                    }
                }
                memcpy(p + 24 + k, &ref, 4);
            }
        }
    }
#else
    Collapsed c;
    rc = collapse(s, c);
    if (rc) return rc;

    // ---- nodes: (lmin.x lmax.x rmin.x rmax.x) (..y) (..z) (left_ref right_ref 0 0)
    float* nodes = reinterpret_cast<float*>(base + info->off_nodes);
    auto child = [&](int ref_node, float* bmin, float* bmax) -> int {
        if (ref_node < 0) { for (int k = 0; k < 3; k++) { bmin[k] = 0.0f; bmax[k] = 0.0f; } return kRefNone; }
        for (int k = 0; k < 3; k++) { bmin[k] = s->bvh_bbox_min[3 * ref_node + k]; bmax[k] = s->bvh_bbox_max[3 * ref_node + k]; }
        if (s->bvh_prim_idx[ref_node] >= 0) {
            const int pt = s->bvh_prim_type[ref_node], pi = s->bvh_prim_idx[ref_node];
            return leaf_ref(pt, pi, shading_class(s->mat[pt], pi));
        }
        return c.new_index[ref_node];
    };
    // Every packed node has two real children, so the kernels' slab test needs no "child exists"
    // check: a node with a single child (only a one-primitive scene's root in the reference's trees)
    // references that child twice -- the second visit finds nothing closer (strict <, kernels.py:691).
    auto write_node = [&](int k, int lnode, int rnode) {
        if (lnode < 0) lnode = rnode;
        if (rnode < 0) rnode = lnode;
        float lmin[3], lmax[3], rmin[3], rmax[3];
        int lref = child(lnode, lmin, lmax), rref = child(rnode, rmin, rmax);
        float* p = nodes + 16 * (size_t)k;
        for (int a = 0; a < 3; a++) { p[4 * a] = lmin[a]; p[4 * a + 1] = lmax[a]; p[4 * a + 2] = rmin[a]; p[4 * a + 3] = rmax[a]; }
        memcpy(p + 12, &lref, 4);
        memcpy(p + 13, &rref, 4);
    };
    if (c.root_is_leaf) {
        write_node(0, 0, -1);
    } else {
        for (size_t k = 0; k < c.order.size(); k++) write_node((int)k, s->bvh_left[c.order[k]], s->bvh_right[c.order[k]]);
    }
#endif

    // ---- primitives
    if (s->num_spheres) memcpy(base + info->off_spheres, s->sphere_data, (size_t)s->num_spheres * 16);
    float* q = reinterpret_cast<float*>(base + info->off_quads);
    for (int i = 0; i < s->num_quads; i++, q += 16) {
        const float *Q = s->quad_Q + 3 * i, *u = s->quad_u + 3 * i, *v = s->quad_v + 3 * i, *n = s->quad_normal + 3 * i,
                    *w = s->quad_w + 3 * i;
        q[0] = n[0]; q[1] = n[1]; q[2] = n[2]; q[3] = s->quad_D[i];
        q[4] = Q[0]; q[5] = Q[1]; q[6] = Q[2]; q[7] = w[0];
        q[8] = u[0]; q[9] = u[1]; q[10] = u[2]; q[11] = w[1];
        q[12] = v[0]; q[13] = v[1]; q[14] = v[2]; q[15] = w[2];
    }
    float* t = reinterpret_cast<float*>(base + info->off_tris);
    for (int i = 0; i < s->num_triangles; i++, t += 12) {
        const float *v0 = s->tri_v0 + 3 * i, *e1 = s->tri_edge1 + 3 * i, *e2 = s->tri_edge2 + 3 * i, *n = s->tri_normal + 3 * i;
        t[0] = v0[0]; t[1] = v0[1]; t[2] = v0[2]; t[3] = n[0];
        t[4] = e1[0]; t[5] = e1[1]; t[6] = e1[2]; t[7] = n[1];
        t[8] = e2[0]; t[9] = e2[1]; t[10] = e2[2]; t[11] = n[2];
    }

    if (info->num_lights > 0) {
        std::vector<LightRec> lights;
        build_lights(s, lights);
        if ((int)lights.size() != info->num_lights) { g_host_error = "pt_pack_scene: info does not match scene"; return PT_ERR_INVALID_ARG; }
        float* lp = reinterpret_cast<float*>(base + info->off_lights);
        for (size_t i = 0; i < lights.size(); i++) {
            memcpy(lp + 8 * i, &lights[i].ref, 4);
            lp[8 * i + 1] = lights[i].area;
            lp[8 * i + 2] = lights[i].emit[0]; lp[8 * i + 3] = lights[i].emit[1]; lp[8 * i + 4] = lights[i].emit[2];
        }
    }
    if (info->full_features && s->sphere_motion) {
        float* mo = reinterpret_cast<float*>(base + info->off_sphere_motion);
        for (int i = 0; i < s->num_spheres; i++) { mo[4 * i] = s->sphere_motion[3 * i]; mo[4 * i + 1] = s->sphere_motion[3 * i + 1]; mo[4 * i + 2] = s->sphere_motion[3 * i + 2]; }
    }

    // ---- materials
    MatTable mt;
    build_materials(s, mt);
    if ((int)mt.recs.size() != info->num_materials) { g_host_error = "pt_pack_scene: info does not match scene"; return PT_ERR_INVALID_ARG; }
    if (!mt.ids.empty()) memcpy(base + info->off_prim_mat, mt.ids.data(), mt.ids.size() * 4);
    if (!mt.recs.empty()) memcpy(base + info->off_materials, mt.recs.data(), mt.recs.size() * 48);

    // ---- Perlin tables
    float* pv = reinterpret_cast<float*>(base + info->off_perlin);
    uint8_t* pp = reinterpret_cast<uint8_t*>(base + info->off_perlin + 256 * 16);
    for (int i = 0; i < 256; i++) {
        if (s->perlin_randvec) { pv[4 * i] = s->perlin_randvec[3 * i]; pv[4 * i + 1] = s->perlin_randvec[3 * i + 1]; pv[4 * i + 2] = s->perlin_randvec[3 * i + 2]; }
        if (s->perm_x) pp[i] = (uint8_t)s->perm_x[i];
        if (s->perm_y) pp[256 + i] = (uint8_t)s->perm_y[i];
        if (s->perm_z) pp[512 + i] = (uint8_t)s->perm_z[i];
    }

    // ---- image textures: RGB8 -> RGBA8
    for (int i = 0; i < s->num_images; i++) {
        uint8_t* dst = reinterpret_cast<uint8_t*>(base + info->off_images + info->image_off[i]);
        const uint8_t* src = s->images[i];
        const size_t npx = (size_t)s->image_w[i] * s->image_h[i];
        for (size_t k = 0; k < npx; k++) { dst[4 * k] = src[3 * k]; dst[4 * k + 1] = src[3 * k + 1]; dst[4 * k + 2] = src[3 * k + 2]; dst[4 * k + 3] = 255; }
    }
    return PT_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Native binned-SAH builder.  Reproduces SAHBVHBuilder (sah_bvh_builder.py:167-418) decision for
// decision in float32 -- the arithmetic numpy >= 2 performs on the reference's float32 arrays with
// weak Python scalars -- so the flattened arrays are bit-identical to the reference's.
// ------------------------------------------------------------------------------------------------
namespace {

struct BNode {
    float bmin[3], bmax[3];
    int left = -1, right = -1;
    int prim_type = -1, prim_idx = -1;
};

struct Builder {
    int n;
    const float *pbmin, *pbmax, *pc;
    const int32_t *ptype, *pidx;
    std::vector<BNode> nodes;

    static float surface_area(const float* mn, const float* mx) {       // :26-29
        float e0 = mx[0] - mn[0], e1 = mx[1] - mn[1], e2 = mx[2] - mn[2];
        return 2.0f * ((e0 * e1 + e1 * e2) + e2 * e0);
    }

    // _find_best_split, :243-336
    void find_best_split(const std::vector<int>& prims, const float* pmin, const float* pmax, int& best_axis, float& best_pos) {
        constexpr int NB = 16;
        const float inf = std::numeric_limits<float>::infinity();
        best_axis = 0;
        int best_bucket = 0;
        float best_cost = inf;
        bool found = false;
        for (int axis = 0; axis < 3; axis++) {
            float mn = inf, mx = -inf;
            for (int p : prims) { float c = pc[3 * p + axis]; if (c < mn) mn = c; if (c > mx) mx = c; }
            if (mx - mn < 1e-10f) continue;
            int cnt[NB];
            float bmn[NB][3], bmx[NB][3];
            for (int b = 0; b < NB; b++) { cnt[b] = 0; for (int k = 0; k < 3; k++) { bmn[b][k] = inf; bmx[b][k] = -inf; } }
            const float extent = mx - mn;
            for (int p : prims) {
                float off = (pc[3 * p + axis] - mn) / extent;
                int b = (int)(off * 16.0f);
                if (b > NB - 1) b = NB - 1;
                cnt[b]++;
                for (int k = 0; k < 3; k++) { bmn[b][k] = std::min(bmn[b][k], pbmin[3 * p + k]); bmx[b][k] = std::max(bmx[b][k], pbmax[3 * p + k]); }
            }
            int lc[NB - 1], rc[NB - 1];
            float lmn[NB - 1][3], lmx[NB - 1][3], rmn[NB - 1][3], rmx[NB - 1][3];
            {
                float a[3] = {inf, inf, inf}, b3[3] = {-inf, -inf, -inf};
                int run = 0;
                for (int i = 0; i < NB - 1; i++) {
                    run += cnt[i];
                    for (int k = 0; k < 3; k++) { a[k] = std::min(a[k], bmn[i][k]); b3[k] = std::max(b3[k], bmx[i][k]); lmn[i][k] = a[k]; lmx[i][k] = b3[k]; }
                    lc[i] = run;
                }
            }
            {
                float a[3] = {inf, inf, inf}, b3[3] = {-inf, -inf, -inf};
                int run = 0;
                for (int i = NB - 2; i >= 0; i--) {
                    run += cnt[i + 1];
                    for (int k = 0; k < 3; k++) { a[k] = std::min(a[k], bmn[i + 1][k]); b3[k] = std::max(b3[k], bmx[i + 1][k]); rmn[i][k] = a[k]; rmx[i][k] = b3[k]; }
                    rc[i] = run;
                }
            }
            const float parent_sa = surface_area(pmin, pmax);
            for (int i = 0; i < NB - 1; i++) {
                if (lc[i] == 0 || rc[i] == 0) continue;
                float lsa = surface_area(lmn[i], lmx[i]), rsa = surface_area(rmn[i], rmx[i]);
                float cost = (1.0f + (lsa / parent_sa) * 1.5f * (float)lc[i]) + (rsa / parent_sa) * 1.5f * (float)rc[i];
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bucket = i; found = true; }
            }
        }
        if (found) {
            float mn = inf, mx = -inf;
            for (int p : prims) { float c = pc[3 * p + best_axis]; if (c < mn) mn = c; if (c > mx) mx = c; }
            float extent = mx - mn;
            best_pos = mn + ((float)(best_bucket + 1) * extent) / 16.0f;
        } else {
            std::vector<int> sorted(prims);
            std::stable_sort(sorted.begin(), sorted.end(), [&](int a, int b) { return pc[3 * a] < pc[3 * b]; });
            best_pos = pc[3 * sorted[sorted.size() / 2]];
            best_axis = 0;
        }
    }

    // _build_recursive, :179-241
    int build(std::vector<int>& prims) {
        const float inf = std::numeric_limits<float>::infinity();
        const int me = (int)nodes.size();
        nodes.emplace_back();
        {
            BNode& nd = nodes[me];
            for (int k = 0; k < 3; k++) { nd.bmin[k] = inf; nd.bmax[k] = -inf; }
            for (int p : prims)
                for (int k = 0; k < 3; k++) { nd.bmin[k] = std::min(nd.bmin[k], pbmin[3 * p + k]); nd.bmax[k] = std::max(nd.bmax[k], pbmax[3 * p + k]); }
            if (prims.size() == 1) { nd.prim_type = ptype[prims[0]]; nd.prim_idx = pidx[prims[0]]; return me; }
        }
        int axis;
        float pos;
        float pmin[3], pmax[3];
        memcpy(pmin, nodes[me].bmin, sizeof pmin);
        memcpy(pmax, nodes[me].bmax, sizeof pmax);
        find_best_split(prims, pmin, pmax, axis, pos);
        std::vector<int> lp, rp;
        for (int p : prims) (pc[3 * p + axis] < pos ? lp : rp).push_back(p);
        if (lp.empty() || rp.empty()) {
            std::vector<int> sorted(prims);
            std::stable_sort(sorted.begin(), sorted.end(), [&](int a, int b) { return pc[3 * a + axis] < pc[3 * b + axis]; });
            size_t mid = sorted.size() / 2;
            lp.assign(sorted.begin(), sorted.begin() + mid);
            rp.assign(sorted.begin() + mid, sorted.end());
        }
        std::vector<int>().swap(prims);   // free before recursing
        int l = build(lp);
        int r = build(rp);
        nodes[me].left = l;
        nodes[me].right = r;
        return me;
    }
};

}  // namespace

extern "C" int pt_bvh_build_sah(int32_t n_prims, const float* prim_bbox_min, const float* prim_bbox_max,
                                const float* prim_centroid, const int32_t* prim_type, const int32_t* prim_idx,
                                float* out_bbox_min, float* out_bbox_max, int32_t* out_left, int32_t* out_right,
                                int32_t* out_parent, int32_t* out_prim_type, int32_t* out_prim_idx) {
    if (n_prims < 0) { g_host_error = "pt_bvh_build_sah: negative primitive count"; return PT_ERR_INVALID_ARG; }
    if (n_prims == 0) return 0;
    if (!prim_bbox_min || !prim_bbox_max || !prim_centroid || !prim_type || !prim_idx || !out_bbox_min || !out_bbox_max ||
        !out_left || !out_right || !out_parent || !out_prim_type || !out_prim_idx) {
        g_host_error = "pt_bvh_build_sah: NULL argument";
        return PT_ERR_INVALID_ARG;
    }
    Builder b{n_prims, prim_bbox_min, prim_bbox_max, prim_centroid, prim_type, prim_idx, {}};
    b.nodes.reserve((size_t)2 * n_prims);
    std::vector<int> all(n_prims);
    for (int i = 0; i < n_prims; i++) all[i] = i;
    // The recursion creates nodes in depth-first pre-order (node, left subtree, right subtree):
    // exactly the order flatten() assigns indices in (:359-388), so no second pass is needed.
    b.build(all);
    const int n = (int)b.nodes.size();
    for (int i = 0; i < n; i++) out_parent[i] = -1;
    for (int i = 0; i < n; i++) {
        const BNode& nd = b.nodes[i];
        for (int k = 0; k < 3; k++) { out_bbox_min[3 * i + k] = nd.bmin[k]; out_bbox_max[3 * i + k] = nd.bmax[k]; }
        out_left[i] = nd.left;
        out_right[i] = nd.right;
        out_prim_type[i] = nd.prim_type;
        out_prim_idx[i] = nd.prim_idx;
        if (nd.left >= 0) out_parent[nd.left] = i;
        if (nd.right >= 0) out_parent[nd.right] = i;
    }
    return n;
}
