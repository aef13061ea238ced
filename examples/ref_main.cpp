// b200cd_run — the reference's main() (reference CollisionDetection/main.cu:47-174) with its body replaced by calls
// into libb200cd.so. This is the binding a maintainer of the reference adds, compiled by
// __graft_entry__.build() and run by tests/test_gpu_example.py, so the C ABI is exercised from C++ and not only
// through ctypes. Plain C++ (g++), no CUDA in this file: everything GPU-side is behind include/b200cd.h.
//
//     b200cd_run mesh.obj [--validate] [--auto-box] [--device N]
//
// Prints what main.cu:117-154 prints: mesh statistics (load_obj.h:117-119), optionally the structural self-check
// counters (main.cu:119,127), the contact count (main.cu:147), one "%07u - %07u" line per colliding pair, lower
// triangle ID first (main.cu:151), and the sorted set of colliding triangle IDs (makeAndPrintSet, main.cu:33-45) -
// the latter computed on the device by b200cd_unique_triangles instead of a host std::set.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "b200cd.h"

static int fail(b200cd_ctx* ctx, const char* what, int rc) {
    // the reference prints and exit()s (common/book.h:21-31, load_obj.h:34,60,73); here every failure is a status
    fprintf(stderr, "b200cd_run: %s: %s%s%s\n", what, b200cd_strerror(rc), ctx ? " - " : "", ctx ? b200cd_last_error(ctx) : "");
    return 1;
}

int main(int argc, char** argv) {
    const char* path = nullptr;
    bool validate = false, auto_box = false;
    int device = 0;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--validate")) validate = true;
        else if (!strcmp(argv[i], "--auto-box")) auto_box = true;
        else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
        else path = argv[i];
    }
    if (!path) {
        fprintf(stderr, "usage: b200cd_run mesh.obj [--validate] [--auto-box] [--device N]\n");
        return 2;
    }
    b200cd_ctx* ctx = nullptr;
    b200cd_mesh* mesh = nullptr;
    b200cd_bvh* bvh = nullptr;
    int rc = b200cd_create(device, &ctx);                                   // was: implicit device 0
    if (rc != B200CD_OK) return fail(nullptr, "b200cd_create", rc);

    rc = b200cd_mesh_load_obj(ctx, path, &mesh);                            // main.cu:64 loadObj(...)
    if (rc != B200CD_OK) return fail(ctx, "b200cd_mesh_load_obj", rc);
    uint32_t nverts = 0, ntris = 0;
    b200cd_mesh_info(mesh, &nverts, &ntris);
    printf("\nObj File Loaded:\n- %u vertexes loaded\n- %u triangles loaded\n", nverts, ntris);   // load_obj.h:117-119

    b200cd_params p;
    b200cd_default_params(&p);                                              // morton.h:45,51,57
    p.auto_box = auto_box ? 1 : 0;
    rc = b200cd_bvh_build(ctx, mesh, &p, &bvh);                             // main.cu:78-107 (+ load_obj.h:89-107)
    if (rc != B200CD_OK) return fail(ctx, "b200cd_bvh_build", rc);

    if (validate) {                                                         // main.cu:113-136
        b200cd_checks c;
        rc = b200cd_bvh_validate(ctx, bvh, mesh, &c);
        if (rc != B200CD_OK) return fail(ctx, "b200cd_bvh_validate", rc);
        printf("Internal node check result: nullParentnum = %u, wrongBoundCount=%u, nullChildCount=%u, notInternalCount=0, uninitBoxCount=%u\n",
               c.null_parent_internal, c.wrong_bound_count, c.null_child, c.uninit_box_internal);
        printf("Leaf node check result: nullParentnum = %u, nullTriangle=%u, notLeafCount=0, illegalBoxCount=%u\n",
               c.null_parent_leaf, c.bad_triangle, c.uninit_box_leaf);
        printf("wrong morton sort count: %u\n", c.unsorted_keys);          // load_obj.h:116
    }

    std::vector<uint32_t> pairs(2 * 500);                                   // main.cu:74,81: the reference's fixed 500 pairs
    uint64_t count = 0;
    rc = b200cd_self_collide(ctx, bvh, pairs.data(), pairs.size() / 2, &count, /*sorted*/ 1);   // main.cu:142-146
    if (rc == B200CD_E_CAPACITY) {                                          // the reference would have overrun its buffer
        pairs.resize(2 * count);
        rc = b200cd_self_collide(ctx, bvh, pairs.data(), count, &count, 1);
    }
    if (rc != B200CD_OK) return fail(ctx, "b200cd_self_collide", rc);

    b200cd_stats st;
    b200cd_get_stats(ctx, &st);                                             // printElapsedTime, main.cu:19-24
    printf("Time of build (morton+sort+hierarchy+boxes): %.3f ms\n", st.ms_build);
    printf("Time of findCollisions: %.3f ms\n", st.ms_query);

    printf("\n\n- contact val = %llu\n", (unsigned long long)count);       // main.cu:147
    printf("\nCollision pair (%llu triangle pairs in total):\n", (unsigned long long)count);   // main.cu:149
    for (uint64_t i = 0; i < count; ++i) printf("%07u - %07u\n", pairs[2 * i], pairs[2 * i + 1]);   // main.cu:151

    uint64_t nids = 0;                                                      // makeAndPrintSet, main.cu:33-45,154
    rc = b200cd_unique_triangles(ctx, bvh, nullptr, 0, &nids);
    if (rc != B200CD_OK && rc != B200CD_E_CAPACITY) return fail(ctx, "b200cd_unique_triangles", rc);
    std::vector<uint32_t> ids(nids ? nids : 1);
    if (nids) {
        rc = b200cd_unique_triangles(ctx, bvh, ids.data(), nids, &nids);
        if (rc != B200CD_OK) return fail(ctx, "b200cd_unique_triangles", rc);
    }
    printf("\n\nCollision Triangles:(%llu points in total):\n", (unsigned long long)nids);   // main.cu:40
    for (uint64_t i = 0; i < nids; ++i) printf("%u\n", ids[i]);
    printf("- Successfully Return\n");                                      // main.cu:172

    b200cd_bvh_destroy(bvh);
    b200cd_mesh_destroy(mesh);
    b200cd_destroy(ctx);
    return 0;
}
