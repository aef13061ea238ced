/*
 * ref_driver.cu — thin driver around the REFERENCE'S OWN host functions.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md). Built by oracle/Makefile into
 * oracle/_ref/libref_cd.so, *including the reference headers from where they lie*
 * (-I/root/reference/CollisionDetection); no reference source is copied into this
 * repository. It exists to (1) pin the CPU restatement oracle/cd_oracle.c against
 * the real reference code, (2) generate tests/golden/ fixtures, (3) serve as the
 * "reference" CPU baseline bench.py times on the GPU box's host cores.
 *
 * What is the reference's and what is ours:
 *   reference (called unmodified): loadObj (load_obj.h:24), morton3D (morton.h:70),
 *     thrust::sort_by_key as load_obj.h:107 uses it, fillLeafNodesCpu (cpu.cuh:89),
 *     generateHierarchyParallelCpu (cpu.cuh:110), calBoundingBoxCpu (cpu.cuh:167),
 *     findCollisionIterativeCpu (cpu.cuh:196), determineRangeCpu / findSplitCpu
 *     (cpu.cuh:65,22), checkTriangleContact (tri_contact.cuh:19), checkBoxOverlap
 *     (box.cuh:40).
 *   ours: this driver (the reference's cpu_main.cu is entirely commented out,
 *     cpu_main.cu:1-117); clzll() — the reference's cpu_math.cpp:12-27 is broken
 *     (never terminates for 0, wrong mask), so we supply device-__clzll semantics,
 *     which is what the reference's GPU path uses (bvh.cuh:48); zeroed node storage
 *     (Node() leaves isLeaf/idx/childCount uninitialised, bvh.cuh:38-42); a pair
 *     buffer that grows (main.cu:81 fixes it at 500 pairs).
 *   also here (ref_gpu_run): the reference's own GPU KERNELS - fillLeafNodes (bvh.cuh:125),
 *     generateHierarchyParallel (bvh.cuh:146), calBoundingBox (bvh.cuh:258), findCollisions
 *     (collision.cuh:73) - launched with the grids of main.cu:92,99,107,140-142 on whatever GPU is
 *     present (sm_100a SASS), timed with CUDA events like main.cu:91-94. bench.py prints them as
 *     `gpu_reference_baseline` next to our numbers for the flag-sized mesh.
 *   not linked: findCollisionsCpu (cpu.cuh:247-271) reads threadIdx in host code;
 *     it is dropped by -fvisibility=hidden + --gc-sections.
 */
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "load_obj.h"
#include "cpu.cuh"
#include "collision.cuh"   // the reference's GPU kernels (findCollisions; bvh.cuh's kernels come with cpu.cuh)

int clzll(unsigned long long x, int) { return x ? __builtin_clzll(x) : 64; }

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {

struct RefMesh {
    std::vector<vec3f> verts;
    std::vector<Triangle> tris;   // sorted by Morton code after load (load_obj.h:107)
    std::vector<unsigned long long> mortons;
    Node* leaves = nullptr;
    Node* inner = nullptr;
    unsigned int wrong_parent = 0;
    std::vector<unsigned int> pairs;
    unsigned int npairs = 0;
    double ms_load = 0, ms_fill = 0, ms_hier = 0, ms_refit = 0, ms_query = 0;
    ~RefMesh() { free(leaves); free(inner); }
};

double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

}  // namespace

REF_API void* ref_load_obj(const char* path) {
    auto* m = new RefMesh;
    double t0 = now_ms();
    loadObj(std::string(path), m->verts, m->tris, m->mortons);
    m->ms_load = now_ms() - t0;
    return m;
}

/* Same steps as the face branch of loadObj (load_obj.h:76-107) — push vertices,
 * then for each face compute centroid + morton3D with the reference's own
 * function, then the reference's thrust::sort_by_key call — but fed from arrays
 * so multi-million-triangle meshes do not need to go through OBJ text. */
REF_API void* ref_from_arrays(const float* xyz, uint32_t nverts, const uint32_t* idx, uint32_t ntris) {
    auto* m = new RefMesh;
    double t0 = now_ms();
    m->verts.reserve(nverts);
    for (uint32_t i = 0; i < nverts; ++i) m->verts.push_back(vec3f(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
    m->tris.reserve(ntris);
    m->mortons.reserve(ntris);
    for (uint32_t t = 0; t < ntris; ++t) {
        Triangle f;
        f.vIdx[0] = idx[3 * t]; f.vIdx[1] = idx[3 * t + 1]; f.vIdx[2] = idx[3 * t + 2];
        vec3f *p1 = &m->verts[f.vIdx[0]], *p2 = &m->verts[f.vIdx[1]], *p3 = &m->verts[f.vIdx[2]];
        double xAvg = (p1->x + p2->x + p3->x) / 3, yAvg = (p1->y + p2->y + p3->y) / 3,
               zAvg = (p1->z + p2->z + p3->z) / 3;
        f.ID = (unsigned int)m->tris.size();
        f.morton = morton3D(xAvg, yAvg, zAvg);
        f.ax = xAvg; f.ay = yAvg; f.az = zAvg;
        m->tris.push_back(f);
        m->mortons.push_back(f.morton);
    }
    thrust::sort_by_key(m->mortons.begin(), m->mortons.end(), m->tris.begin());
    m->ms_load = now_ms() - t0;
    return m;
}

/* Same as ref_from_arrays, but the sort keys come from the caller instead of morton3D, whose
 * normalisation is hard-coded for the flag mesh (morton.h:43-58): a unit-cube mesh has x below the
 * box origin (UB in morton.h:80) and, at 16-64 M triangles, duplicate 60-bit codes, for which the
 * reference builds a malformed tree (no tie-break in cpu_math.h:10 / bvh.cuh:48). The keys must be
 * unique; ANY strictly increasing key sequence gives the reference's hierarchy a valid binary tree
 * over the sorted leaf order, and the emitted pair SET does not depend on the tree.
 * Everything after the key - thrust::sort_by_key, fillLeafNodesCpu, generateHierarchyParallelCpu,
 * calBoundingBoxCpu, findCollisionIterativeCpu, checkTriangleContactHelper - is the reference's.
 * Used by tests/golden/make_checksums.py for the full-size headline workloads. */
REF_API void* ref_from_arrays_keyed(const float* xyz, uint32_t nverts, const uint32_t* idx, uint32_t ntris,
                                    const uint64_t* keys) {
    auto* m = new RefMesh;
    double t0 = now_ms();
    m->verts.reserve(nverts);
    for (uint32_t i = 0; i < nverts; ++i) m->verts.push_back(vec3f(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
    m->tris.reserve(ntris);
    m->mortons.reserve(ntris);
    for (uint32_t t = 0; t < ntris; ++t) {
        Triangle f;
        f.vIdx[0] = idx[3 * t]; f.vIdx[1] = idx[3 * t + 1]; f.vIdx[2] = idx[3 * t + 2];
        vec3f *p1 = &m->verts[f.vIdx[0]], *p2 = &m->verts[f.vIdx[1]], *p3 = &m->verts[f.vIdx[2]];
        f.ID = t;
        f.morton = keys[t];
        f.ax = (p1->x + p2->x + p3->x) / 3; f.ay = (p1->y + p2->y + p3->y) / 3; f.az = (p1->z + p2->z + p3->z) / 3;
        m->tris.push_back(f);
        m->mortons.push_back(f.morton);
    }
    thrust::sort_by_key(m->mortons.begin(), m->mortons.end(), m->tris.begin());
    m->ms_load = now_ms() - t0;
    return m;
}

REF_API void ref_free(void* h) { delete static_cast<RefMesh*>(h); }
REF_API uint32_t ref_num_verts(void* h) { return (uint32_t) static_cast<RefMesh*>(h)->verts.size(); }
REF_API uint32_t ref_num_tris(void* h) { return (uint32_t) static_cast<RefMesh*>(h)->tris.size(); }

/* vertices as the floats they were parsed as; triangle indices in ID (face) order */
REF_API void ref_get_mesh(void* h, float* xyz, uint32_t* idx) {
    auto* m = static_cast<RefMesh*>(h);
    for (size_t i = 0; i < m->verts.size(); ++i) {
        xyz[3 * i] = (float)m->verts[i].x; xyz[3 * i + 1] = (float)m->verts[i].y; xyz[3 * i + 2] = (float)m->verts[i].z;
    }
    for (const Triangle& t : m->tris)
        for (int k = 0; k < 3; ++k) idx[3 * (size_t)t.ID + k] = t.vIdx[k];
}

REF_API void ref_get_sorted(void* h, uint64_t* keys, uint32_t* ids) {
    auto* m = static_cast<RefMesh*>(h);
    for (size_t i = 0; i < m->tris.size(); ++i) { keys[i] = m->mortons[i]; ids[i] = m->tris[i].ID; }
}

/* cpu_main.cu:55-90 equivalent */
REF_API unsigned int ref_build(void* h) {
    auto* m = static_cast<RefMesh*>(h);
    const int n = (int)m->tris.size();
    free(m->leaves); free(m->inner);
    m->leaves = static_cast<Node*>(calloc((size_t)n, sizeof(Node)));
    m->inner = static_cast<Node*>(calloc((size_t)(n > 1 ? n - 1 : 1), sizeof(Node)));
    m->wrong_parent = 0;
    double t0 = now_ms();
    fillLeafNodesCpu(m->tris.data(), n, m->leaves);
    double t1 = now_ms();
    generateHierarchyParallelCpu(m->mortons.data(), n, m->leaves, m->inner, &m->wrong_parent);
    double t2 = now_ms();
    calBoundingBoxCpu(m->leaves, m->verts.data(), (unsigned int)n);
    double t3 = now_ms();
    m->ms_fill = t1 - t0; m->ms_hier = t2 - t1; m->ms_refit = t3 - t2;
    return m->wrong_parent;
}

/* unified numbering: internal i -> i, leaf j -> (n-1)+j */
REF_API void ref_get_nodes(void* h, int32_t* left, int32_t* right, int32_t* parent, uint32_t* bounded,
                           double* bounds) {
    auto* m = static_cast<RefMesh*>(h);
    const int n = (int)m->tris.size();
    auto number = [&](Node* p) -> int32_t {
        if (!p) return -1;
        if (p >= m->leaves && p < m->leaves + n) return (int32_t)((n - 1) + (p - m->leaves));
        return (int32_t)(p - m->inner);
    };
    for (int i = 0; i < 2 * n - 1; ++i) {
        Node* nd = i < n - 1 ? &m->inner[i] : &m->leaves[i - (n - 1)];
        if (i < n - 1) { left[i] = number(nd->childA); right[i] = number(nd->childB); bounded[i] = nd->bounded; }
        parent[i] = number(nd->parent);
        double* b = bounds + 6 * (size_t)i;
        b[0] = nd->box.x1; b[1] = nd->box.y1; b[2] = nd->box.z1;
        b[3] = nd->box.x2; b[4] = nd->box.y2; b[5] = nd->box.z2;
    }
}

/* the loop of findCollisionsCpu (cpu.cuh:268-270) around the reference's
 * findCollisionIterativeCpu, with a pair buffer that cannot overflow */
REF_API uint64_t ref_collide(void* h) {
    auto* m = static_cast<RefMesh*>(h);
    const int n = (int)m->tris.size();
    m->npairs = 0;
    m->pairs.assign(1 << 20, 0u);
    double t0 = now_ms();
    for (int i = 0; i < n; ++i) {
        if (m->pairs.size() / 2 - m->npairs < (size_t)n + 16) m->pairs.resize(m->pairs.size() * 2 + 2 * (size_t)n);
        findCollisionIterativeCpu(&m->inner[0], m->leaves[i].triangle, &m->leaves[i].box, m->verts.data(),
                                  &m->npairs, m->pairs.data());
    }
    m->ms_query = now_ms() - t0;
    return m->npairs;
}

/* Same reference function per query, but the (independent) queries are split over host
 * threads — OUR parallelisation, used only for bench.py's "all host threads" CPU arm. Each
 * thread appends to its own buffer through the reference's count/out arguments; the lists are
 * concatenated afterwards. The reference itself is serial (cpu.cuh:268). */
REF_API uint64_t ref_collide_mt(void* h, int nthreads) {
    auto* m = static_cast<RefMesh*>(h);
    const int n = (int)m->tris.size();
    if (nthreads < 1) nthreads = 1;
    std::vector<std::vector<unsigned int>> bufs(nthreads);
    std::vector<unsigned int> counts(nthreads, 0u);
    double t0 = now_ms();
    auto work = [&](int t) {
        std::vector<unsigned int>& buf = bufs[t];
        buf.assign(1 << 16, 0u);
        unsigned int cnt = 0;
        // interleaved blocks of 1024 sorted leaves: neighbouring queries cost about the same
        for (int blk = t; blk * 1024 < n; blk += nthreads) {
            const int lo = blk * 1024, hi = lo + 1024 < n ? lo + 1024 : n;
            for (int i = lo; i < hi; ++i) {
                if (buf.size() / 2 - cnt < (size_t)n + 16) buf.resize(buf.size() * 2 + 2 * (size_t)n);
                findCollisionIterativeCpu(&m->inner[0], m->leaves[i].triangle, &m->leaves[i].box, m->verts.data(),
                                          &cnt, buf.data());
            }
        }
        counts[t] = cnt;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    m->npairs = 0;
    for (int t = 0; t < nthreads; ++t) m->npairs += counts[t];
    m->pairs.assign(2 * (size_t)m->npairs + 2, 0u);
    size_t off = 0;
    for (int t = 0; t < nthreads; ++t) {
        memcpy(m->pairs.data() + off, bufs[t].data(), (size_t)counts[t] * 8);
        off += 2 * (size_t)counts[t];
    }
    m->ms_query = now_ms() - t0;
    return m->npairs;
}

REF_API void ref_get_pairs(void* h, uint32_t* out) {
    auto* m = static_cast<RefMesh*>(h);
    memcpy(out, m->pairs.data(), (size_t)m->npairs * 8);
}

REF_API void ref_get_timing(void* h, double* ms5) {
    auto* m = static_cast<RefMesh*>(h);
    ms5[0] = m->ms_load; ms5[1] = m->ms_fill; ms5[2] = m->ms_hier; ms5[3] = m->ms_refit; ms5[4] = m->ms_query;
}

/* ---- the reference's GPU path (main.cu:78-146) on the mesh held by `h` ----
 * Ours: zeroed node arrays and pair counter (main.cu:82-85 cudaMalloc only: the reference relies on fresh
 * allocations being zero), a pair buffer of `pair_cap` pairs instead of 500 (main.cu:81; the
 * kernel appends unguarded, collision.cuh:40-42, so pair_cap must exceed the true count), `repeats` timed runs
 * (stage times = minimum over the runs). Returns 0, or a negative code: -1 no device, -2 CUDA error,
 * -3 more pairs than pair_cap. ms4 = {fillLeafNodes, generateHierarchyParallel, calBoundingBox, findCollisions}. */
REF_API int ref_gpu_run(void* h, uint32_t pair_cap, int repeats, float* ms4, uint64_t* count_out) {
    auto* m = static_cast<RefMesh*>(h);
    const size_t n = m->tris.size();
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return -1; }
    if (n < 2) return -2;
    vec3f* v_ptr = nullptr; Triangle* t_ptr = nullptr; unsigned long long* m_ptr = nullptr;
    Node *leaf_nodes = nullptr, *internal_nodes = nullptr;
    unsigned int *collision_list = nullptr, *test_val = nullptr, *wrong = nullptr;
    cudaEvent_t ev[5];
    bool ok = true;
    auto C = [&](cudaError_t e) { if (e != cudaSuccess) ok = false; };
    for (auto& e : ev) C(cudaEventCreate(&e));
    C(cudaMalloc((void**)&v_ptr, m->verts.size() * sizeof(vec3f)));
    C(cudaMalloc((void**)&t_ptr, n * sizeof(Triangle)));
    C(cudaMalloc((void**)&m_ptr, n * sizeof(unsigned long long)));
    C(cudaMalloc((void**)&collision_list, 2 * (size_t)pair_cap * sizeof(unsigned int)));
    C(cudaMalloc((void**)&test_val, sizeof(unsigned int)));
    C(cudaMalloc((void**)&wrong, 5 * sizeof(unsigned int)));
    C(cudaMalloc((void**)&leaf_nodes, n * sizeof(Node)));
    C(cudaMalloc((void**)&internal_nodes, (n - 1) * sizeof(Node)));
    if (ok) {
        C(cudaMemcpy(v_ptr, m->verts.data(), m->verts.size() * sizeof(vec3f), cudaMemcpyHostToDevice));
        C(cudaMemcpy(t_ptr, m->tris.data(), n * sizeof(Triangle), cudaMemcpyHostToDevice));
        C(cudaMemcpy(m_ptr, m->mortons.data(), n * sizeof(unsigned long long), cudaMemcpyHostToDevice));
    }
    float best[4] = {1e30f, 1e30f, 1e30f, 1e30f};
    unsigned int count = 0;
    for (int r = 0; ok && r < (repeats < 1 ? 1 : repeats); ++r) {
        C(cudaMemset(leaf_nodes, 0, n * sizeof(Node)));
        C(cudaMemset(internal_nodes, 0, (n - 1) * sizeof(Node)));
        C(cudaMemset(test_val, 0, sizeof(unsigned int)));
        C(cudaMemset(wrong, 0, 5 * sizeof(unsigned int)));
        C(cudaEventRecord(ev[0], 0));
        fillLeafNodes<<<128, 128>>>(t_ptr, (int)n, leaf_nodes);                                      // main.cu:92
        C(cudaEventRecord(ev[1], 0));
        generateHierarchyParallel<<<128, 128>>>(m_ptr, (int)n, leaf_nodes, internal_nodes, wrong);   // main.cu:99
        C(cudaEventRecord(ev[2], 0));
        calBoundingBox<<<128, 128>>>(leaf_nodes, v_ptr, (unsigned int)n);                            // main.cu:107
        C(cudaEventRecord(ev[3], 0));
        dim3 blocks(128, 128), threads(128);                                                         // main.cu:140-141
        findCollisions<<<blocks, threads>>>(&internal_nodes[0], leaf_nodes, v_ptr, (unsigned int)n, test_val, collision_list);
        C(cudaEventRecord(ev[4], 0));
        C(cudaEventSynchronize(ev[4]));
        C(cudaGetLastError());
        for (int k = 0; ok && k < 4; ++k) {
            float ms = 0.f;
            C(cudaEventElapsedTime(&ms, ev[k], ev[k + 1]));
            if (ms < best[k]) best[k] = ms;
        }
        C(cudaMemcpy(&count, test_val, sizeof(unsigned int), cudaMemcpyDeviceToHost));
        if (count > pair_cap) break;
    }
    int rc = ok ? 0 : -2;
    if (ok && count > pair_cap) rc = -3;
    if (rc == 0) {
        m->npairs = count;
        m->pairs.assign(2 * (size_t)count + 2, 0u);
        if (count) C(cudaMemcpy(m->pairs.data(), collision_list, 2 * (size_t)count * sizeof(unsigned int), cudaMemcpyDeviceToHost));
        unsigned int w = 0;
        C(cudaMemcpy(&w, wrong, sizeof(unsigned int), cudaMemcpyDeviceToHost));
        m->wrong_parent = w;
        for (int k = 0; k < 4; ++k) ms4[k] = best[k];
        if (count_out) *count_out = count;
        if (!ok) rc = -2;
    }
    cudaFree(v_ptr); cudaFree(t_ptr); cudaFree(m_ptr); cudaFree(collision_list); cudaFree(test_val); cudaFree(wrong);
    cudaFree(leaf_nodes); cudaFree(internal_nodes);
    for (auto& e : ev) cudaEventDestroy(e);
    cudaGetLastError();
    return rc;
}

/* ---- unit hooks on the reference predicates ---- */
REF_API uint64_t ref_morton3D(double x, double y, double z) { return morton3D(x, y, z); }

REF_API int ref_tri_contact(const double* t) {
    vec3f P1(t[0], t[1], t[2]), P2(t[3], t[4], t[5]), P3(t[6], t[7], t[8]);
    vec3f Q1(t[9], t[10], t[11]), Q2(t[12], t[13], t[14]), Q3(t[15], t[16], t[17]);
    return checkTriangleContact(P1, P2, P3, Q1, Q2, Q3);
}

/* boxes as lo xyz, hi xyz */
REF_API int ref_box_overlap(const double* a, const double* b) {
    Box A, B;
    A.x1 = a[0]; A.y1 = a[1]; A.z1 = a[2]; A.x2 = a[3]; A.y2 = a[4]; A.z2 = a[5];
    B.x1 = b[0]; B.y1 = b[1]; B.z1 = b[2]; B.x2 = b[3]; B.y2 = b[4]; B.z2 = b[5];
    return checkBoxOverlap(&A, &B);
}

/* check.cuh:19-27 known-answer hook, host twins */
REF_API void ref_range_split(const uint64_t* keys, int n, int i, int* out3) {
    std::vector<unsigned long long> k(keys, keys + n);
    Range r = determineRangeCpu(k.data(), n, i);
    out3[0] = r.x; out3[1] = r.y;
    out3[2] = findSplitCpu(k.data(), r.x, r.y);
}
