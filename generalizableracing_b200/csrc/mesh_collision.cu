// mesh_collision.cu -- UAV-vs-terrain-mesh collision count by axis ray casts (last item).
//
// Reference path replaced (L = extensions/diff.lab/diff/lab, QD = extensions/diff.lab_tasks/.../quadcopter_diff): the Warp kernel
// check_uav_collision_ray_kernel (L/utils/mesh_tools.py:128-233) and its launcher get_uav_collision_num_ray (:237-295), consumed by
// the STAGE-0 reward term collision_penalty_custom (QD/mdp/rewards.py:226-242: (num_collision > 2).float(), weight -50,
// QD/racing_ctbr_env.py:299-303) and by S/diff_rl/test/test_collider.py:131-141.  Per UAV, for each of the 17 lattice points of its
// box collider (L/utils/__init__.py:19-37; offset (+-0.707 arm, +-0.707 arm, +-0.5 height) rotated by the attitude), up to six
// axis-aligned rays (+x -x +y -y +z -z) are cast against the terrain mesh; the point counts as colliding as soon as the CLOSEST hit
// of one of them is a back face (sign <= 0: the point is inside a closed body).  Without a lattice only the centre is tested and
// every ray that hits overwrites the result (1 for a back face, 0 for a front face) -- the reference's literal behaviour.
//
// wp.mesh_query_ray (third party: NVIDIA Warp, absent here) is restated from its documented contract: closest intersection with
// 0 <= t <= max_t over all faces, two-sided; sign > 0 when the ray hits the front of the face (the side its counter-clockwise normal
// points to), < 0 otherwise.  wp.Mesh keeps a BVH over the faces; so does this file: a binary BVH built on the host (median split of
// the centroids along the widest axis, <= 4 faces per leaf, children adjacent in memory) and traversed on the device with a
// per-thread stack, near child first, pruned by the best hit so far.  One thread per (UAV, lattice point): 65,536 x 17 rays-bundles
// are 1.1 M independent threads; the per-UAV count is a shared-memory / global integer add.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "gr_math.cuh"
#include "../../include/gracing.h"

namespace gr {

constexpr int kBvhStack = 64;

// node = 2 float4: (lo.xyz, as_int: first face of a leaf | index of the left child), (hi.xyz, as_int: face count of a leaf | 0)
struct BvhHit { float t; float sign; };

__device__ __forceinline__ bool slab(const float4 lo, const float4 hi, const V3 o, const V3 inv_d, const float t_best, float& t_near) {
  const float tx0 = (lo.x - o.x) * inv_d.x, tx1 = (hi.x - o.x) * inv_d.x;
  const float ty0 = (lo.y - o.y) * inv_d.y, ty1 = (hi.y - o.y) * inv_d.y;
  const float tz0 = (lo.z - o.z) * inv_d.z, tz1 = (hi.z - o.z) * inv_d.z;
  const float tmin = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), 0.0f));
  const float tmax = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fminf(fmaxf(tz0, tz1), t_best));
  t_near = tmin;
  // (boxes are padded on the host by a relative epsilon, so a hit exactly on a face of a leaf box is not lost to rounding here)
  return tmin <= tmax;
}

// closest hit of the ray o + t d, 0 <= t <= max_t (Moeller-Trumbore per face; det > 0 <=> the ray runs against the face normal)
__device__ __forceinline__ bool mesh_query_ray(const GrMesh& m, const V3 o, const V3 d, const float max_t, BvhHit& hit) {
  const float4* __restrict__ nodes = reinterpret_cast<const float4*>(m.nodes);
  const float4* __restrict__ tris = reinterpret_cast<const float4*>(m.tris);
  const float tiny = 1e-30f;
  const V3 inv_d = v3(1.0f / (fabsf(d.x) < tiny ? copysignf(tiny, d.x) : d.x), 1.0f / (fabsf(d.y) < tiny ? copysignf(tiny, d.y) : d.y),
                      1.0f / (fabsf(d.z) < tiny ? copysignf(tiny, d.z) : d.z));
  int stack[kBvhStack];
  int sp = 0;
  stack[sp++] = 0;
  float best = max_t, sign = 0.0f;
  bool found = false;
  while (sp > 0) {
    const int n = stack[--sp];
    const float4 lo = __ldg(nodes + 2 * n), hi = __ldg(nodes + 2 * n + 1);
    float tn;
    if (!slab(lo, hi, o, inv_d, best, tn)) continue;
    const int count = __float_as_int(hi.w), first = __float_as_int(lo.w);
    if (count > 0) {
      for (int f = first; f < first + count; ++f) {
        const V3 v0 = xyz(__ldg(tris + 3 * f)), e1 = xyz(__ldg(tris + 3 * f + 1)), e2 = xyz(__ldg(tris + 3 * f + 2));
        const V3 p = cross(d, e2);
        const float det = dot(e1, p);
        if (det == 0.0f) continue;
        const float inv = 1.0f / det;
        const V3 tv = o - v0;
        const float u = dot(tv, p) * inv;
        if (u < 0.0f || u > 1.0f) continue;
        const V3 q = cross(tv, e1);
        const float v = dot(d, q) * inv;
        if (v < 0.0f || u + v > 1.0f) continue;
        const float t = dot(e2, q) * inv;
        if (t >= 0.0f && t <= max_t && (t < best || !found)) { best = t; sign = det > 0.0f ? 1.0f : -1.0f; found = true; }
      }
    } else {
      // children first, first + 1: visit the nearer one first (pushed last)
      const int l = first, r = first + 1;
      const float4 llo = __ldg(nodes + 2 * l), lhi = __ldg(nodes + 2 * l + 1), rlo = __ldg(nodes + 2 * r), rhi = __ldg(nodes + 2 * r + 1);
      float tl, tr;
      const bool hl = slab(llo, lhi, o, inv_d, best, tl), hr = slab(rlo, rhi, o, inv_d, best, tr);
      if (hl && hr) {
        if (sp + 2 > kBvhStack) continue;          // (depth bound of the builder: never reached)
        if (tl <= tr) { stack[sp++] = r; stack[sp++] = l; } else { stack[sp++] = l; stack[sp++] = r; }
      } else if (hl) {
        stack[sp++] = l;
      } else if (hr) {
        stack[sp++] = r;
      }
    }
  }
  hit.t = best; hit.sign = sign;
  return found;
}

__device__ __forceinline__ V3 axis_dir(int k) {      // front, back, left, right, up, down (mesh_tools.py:150-155)
  return k == 0 ? v3(1.f, 0.f, 0.f) : k == 1 ? v3(-1.f, 0.f, 0.f) : k == 2 ? v3(0.f, 1.f, 0.f) : k == 3 ? v3(0.f, -1.f, 0.f) : k == 4 ? v3(0.f, 0.f, 1.f) : v3(0.f, 0.f, -1.f);
}

// one thread per (uav, lattice point); num_lattices == 0: one thread per uav, centre point only
__global__ void __launch_bounds__(128) uav_collision_ray_kernel(const GrMesh mesh, const float* __restrict__ pos, const float* __restrict__ quat_wxyz,
                                                                const float* __restrict__ lattices, const int num_lattices, const int num_uav,
                                                                const float max_dist, const float arm_length, const float height,
                                                                int32_t* __restrict__ collision_num) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int P = num_lattices > 0 ? num_lattices : 1;
  if (idx >= (int64_t)num_uav * P) return;
  const int uav = (int)(idx / P), k = (int)(idx - (int64_t)uav * P);
  const V3 centre = v3(pos[3 * uav], pos[3 * uav + 1], pos[3 * uav + 2]);
  BvhHit hit;
  if (num_lattices == 0) {
    int result = 0;
    bool any = false;
    for (int a = 0; a < 6; ++a)
      if (mesh_query_ray(mesh, centre, axis_dir(a), max_dist, hit)) { result = hit.sign <= 0.0f ? 1 : 0; any = true; }
    if (any) collision_num[uav] = result;
    return;
  }
  const Q4 q = Q4{quat_wxyz[4 * uav], quat_wxyz[4 * uav + 1], quat_wxyz[4 * uav + 2], quat_wxyz[4 * uav + 3]};
  const V3 vec = v3(lattices[3 * k] * 0.707f * arm_length, lattices[3 * k + 1] * 0.707f * arm_length, lattices[3 * k + 2] * 0.5f * height);
  const V3 pt = centre + quat_rotate(q, vec);
  for (int a = 0; a < 6; ++a) {
    if (mesh_query_ray(mesh, pt, axis_dir(a), max_dist, hit) && hit.sign <= 0.0f) {
      atomicAdd(collision_num + uav, 1);
      return;
    }
  }
}

// generic batched query (wp.mesh_query_ray for arrays of rays): t (max_t where nothing was hit), sign (0 where nothing was hit)
__global__ void __launch_bounds__(128) mesh_query_rays_kernel(const GrMesh mesh, const float* __restrict__ origins, const float* __restrict__ dirs,
                                                              const int64_t num_rays, const float max_t, float* __restrict__ t_out,
                                                              float* __restrict__ sign_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num_rays) return;
  BvhHit hit;
  const bool found = mesh_query_ray(mesh, v3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]), v3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]), max_t, hit);
  t_out[i] = found ? hit.t : max_t;
  sign_out[i] = found ? hit.sign : 0.0f;
}

}  // namespace gr

#ifndef GR_CPU_EMUL   // host API (the CPU emulation harness in tests/emul includes only the device code)
// =============================================================================================
// host: BVH builder + C ABI
// =============================================================================================
using namespace gr;

namespace {
struct Box { float lo[3], hi[3]; };
inline void grow(Box& b, const float* p) { for (int a = 0; a < 3; ++a) { b.lo[a] = std::min(b.lo[a], p[a]); b.hi[a] = std::max(b.hi[a], p[a]); } }
inline Box empty_box() { return Box{{INFINITY, INFINITY, INFINITY}, {-INFINITY, -INFINITY, -INFINITY}}; }
}  // namespace

extern "C" int64_t gr_mesh_bvh_max_nodes(int32_t num_faces) { return num_faces <= 0 ? 0 : 2 * (int64_t)num_faces + 1; }

extern "C" int gr_mesh_build_bvh(const float* points, const int32_t* indices, int32_t num_points, int32_t num_faces, float* nodes_out,
                                 int64_t max_nodes, float* tris_out, int32_t* face_ids_out, int32_t* num_nodes_out) {
  if (!points || !indices || !nodes_out || !tris_out || !num_nodes_out) return GR_ERR_NULL;
  if (num_points <= 0 || num_faces <= 0 || max_nodes < gr_mesh_bvh_max_nodes(num_faces)) return GR_ERR_SIZE;
  const int kLeaf = 4;
  std::vector<Box> fbox(num_faces);
  std::vector<float> cen((size_t)num_faces * 3);
  for (int f = 0; f < num_faces; ++f) {
    Box b = empty_box();
    for (int c = 0; c < 3; ++c) {
      const int32_t v = indices[3 * f + c];
      if (v < 0 || v >= num_points) return GR_ERR_SIZE;
      grow(b, points + 3 * (size_t)v);
    }
    fbox[f] = b;
    for (int a = 0; a < 3; ++a) cen[3 * (size_t)f + a] = 0.5f * (b.lo[a] + b.hi[a]);
  }
  std::vector<int32_t> order(num_faces);
  for (int f = 0; f < num_faces; ++f) order[f] = f;
  struct Item { int node, begin, end, depth; };
  std::vector<Item> todo;
  int n_nodes = 1;
  todo.push_back({0, 0, num_faces, 0});
  auto put = [&](int node, const Box& b, int first, int count) {
    float* n = nodes_out + 8 * (size_t)node;
    for (int a = 0; a < 3; ++a) {            // pad: the device slab test must not lose a hit that lies exactly on a box face
      const float ext = std::max(std::fabs(b.lo[a]), std::fabs(b.hi[a]));
      const float eps = 1e-5f * std::max(ext, 1.0f);
      n[a] = b.lo[a] - eps;
      n[4 + a] = b.hi[a] + eps;
    }
    std::memcpy(n + 3, &first, 4);
    std::memcpy(n + 7, &count, 4);
  };
  while (!todo.empty()) {
    const Item it = todo.back();
    todo.pop_back();
    Box b = empty_box(), cb = empty_box();
    for (int k = it.begin; k < it.end; ++k) {
      const Box& fb = fbox[order[k]];
      grow(b, fb.lo); grow(b, fb.hi);
      grow(cb, &cen[3 * (size_t)order[k]]);
    }
    const int count = it.end - it.begin;
    if (count <= kLeaf || it.depth >= kBvhStack - 2) {
      put(it.node, b, it.begin, count);
      continue;
    }
    int axis = 0;
    float ext = cb.hi[0] - cb.lo[0];
    for (int a = 1; a < 3; ++a) if (cb.hi[a] - cb.lo[a] > ext) { ext = cb.hi[a] - cb.lo[a]; axis = a; }
    const int mid = it.begin + count / 2;
    std::nth_element(order.begin() + it.begin, order.begin() + mid, order.begin() + it.end,
                     [&](int32_t x, int32_t y) { const float cx = cen[3 * (size_t)x + axis], cy = cen[3 * (size_t)y + axis]; return cx < cy || (cx == cy && x < y); });
    const int left = n_nodes;
    n_nodes += 2;
    if (n_nodes > max_nodes) return GR_ERR_SIZE;
    put(it.node, b, left, 0);
    todo.push_back({left, it.begin, mid, it.depth + 1});
    todo.push_back({left + 1, mid, it.end, it.depth + 1});
  }
  // a leaf of more than kLeaf faces can only come from the depth bound (degenerate input); the device loop handles any count
  for (int k = 0; k < num_faces; ++k) {
    const int f = order[k];
    const float* a = points + 3 * (size_t)indices[3 * f], *b = points + 3 * (size_t)indices[3 * f + 1], *c = points + 3 * (size_t)indices[3 * f + 2];
    float* t = tris_out + 12 * (size_t)k;
    for (int x = 0; x < 3; ++x) { t[x] = a[x]; t[4 + x] = b[x] - a[x]; t[8 + x] = c[x] - a[x]; }
    t[3] = t[7] = t[11] = 0.0f;
    if (face_ids_out) face_ids_out[k] = f;
  }
  *num_nodes_out = n_nodes;
  return GR_OK;
}

static int check_mesh(const GrMesh* m) {
  if (!m || !m->nodes || !m->tris) return GR_ERR_NULL;
  if (m->num_nodes < 1 || m->num_faces < 1) return GR_ERR_SIZE;
  if ((reinterpret_cast<uintptr_t>(m->nodes) | reinterpret_cast<uintptr_t>(m->tris)) & 15u) return GR_ERR_ALIGN;
  return GR_OK;
}

extern "C" int gr_uav_collision_ray(const GrMesh* mesh, const float* uav_position, const float* uav_quat_wxyz, int32_t num_uav, const float* lattices,
                                    int32_t num_lattices, float max_dist, float arm_length, float height, int32_t* collision_num, void* stream) {
  int rc = check_mesh(mesh);
  if (rc != GR_OK) return rc;
  if (!uav_position || !collision_num || (num_lattices > 0 && (!lattices || !uav_quat_wxyz))) return GR_ERR_NULL;
  if (num_uav <= 0 || num_lattices < 0) return GR_ERR_SIZE;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(collision_num, 0, sizeof(int32_t) * (size_t)num_uav, s);      // torch.zeros(num_uav) of the launcher (:266)
  if (e != cudaSuccess) return (int)e;
  const int64_t threads = (int64_t)num_uav * (num_lattices > 0 ? num_lattices : 1);
  uav_collision_ray_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(*mesh, uav_position, uav_quat_wxyz, lattices, num_lattices, num_uav, max_dist,
                                                                             arm_length, height, collision_num);
  return (int)cudaGetLastError();
}

extern "C" int gr_mesh_query_rays(const GrMesh* mesh, const float* origins, const float* dirs, int64_t num_rays, float max_t, float* t_out, float* sign_out,
                                  void* stream) {
  int rc = check_mesh(mesh);
  if (rc != GR_OK) return rc;
  if (!origins || !dirs || !t_out || !sign_out) return GR_ERR_NULL;
  if (num_rays <= 0) return GR_ERR_SIZE;
  mesh_query_rays_kernel<<<(unsigned)((num_rays + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*mesh, origins, dirs, num_rays, max_t, t_out, sign_out);
  return (int)cudaGetLastError();
}
#endif  // GR_CPU_EMUL
