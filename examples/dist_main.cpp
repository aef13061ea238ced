// b200cd_dist_run — the multi-GPU self-collision of include/b200cd.h (b200cd_dist_*) driven from plain C++, one
// PROCESS per rank, no Python, no MPI, no NCCL, no CUDA in this file: the parent forks the ranks, collects each rank's export blob over a
// pipe and hands every rank the whole set (that is all the host language has to do for the library: move
// B200CD_DIST_BLOB_BYTES per rank once), then relays two barriers. The reference has nothing like this (single GPU,
// SURVEY.md section 5); the per-rank body is the reference's main() (main.cu:47-174) with b200cd_dist_step in place
// of the build + findCollisions launches.
//
//     b200cd_dist_run mesh.obj [--ranks W] [--steps K] [--device D]
//
// Rank r uses GPU (r mod number of GPUs) unless --device pins all ranks to one GPU (they then share it: the same
// peer-memory stores, remote atomics and flag barriers, time-sliced). Rank 0 prints what main.cu:147-154 prints: the
// contact count, the "%07u - %07u" pairs (sorted) and the sorted triangle-ID set.
#include <signal.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "b200cd.h"

static bool read_all(int fd, void* buf, size_t n) {
    char* p = static_cast<char*>(buf);
    while (n) {
        const ssize_t k = read(fd, p, n);
        if (k <= 0) return false;
        p += k;
        n -= (size_t)k;
    }
    return true;
}
static bool write_all(int fd, const void* buf, size_t n) {
    const char* p = static_cast<const char*>(buf);
    while (n) {
        const ssize_t k = write(fd, p, n);
        if (k <= 0) return false;
        p += k;
        n -= (size_t)k;
    }
    return true;
}

static int rank_main(int rank, int world, const char* path, int steps, int pinned_device, int up /* to parent */, int down /* from parent */) {
    int ndev = 0;
    if (b200cd_device_count(&ndev) != B200CD_OK) {
        fprintf(stderr, "rank %d: %s\n", rank, b200cd_strerror(B200CD_E_NODEVICE));
        return 1;
    }
    const int device = pinned_device >= 0 ? pinned_device : rank % ndev;
    b200cd_ctx* ctx = nullptr;
    b200cd_mesh* mesh = nullptr;
    b200cd_dist* dist = nullptr;
    auto fail = [&](const char* what, int rc) {
        fprintf(stderr, "rank %d: %s: %s - %s\n", rank, what, b200cd_strerror(rc), ctx ? b200cd_last_error(ctx) : "");
        return 1;
    };
    int rc = b200cd_create(device, &ctx);
    if (rc != B200CD_OK) return fail("b200cd_create", rc);
    rc = b200cd_mesh_load_obj(ctx, path, &mesh);  // every rank holds the whole mesh (load_obj.h:24 dialect)
    if (rc != B200CD_OK) return fail("b200cd_mesh_load_obj", rc);
    uint32_t nverts = 0, ntris = 0;
    b200cd_mesh_info(mesh, &nverts, &ntris);
    rc = b200cd_dist_create(ctx, (uint32_t)rank, (uint32_t)world, ntris, 1.5, 0, &dist);
    if (rc != B200CD_OK) return fail("b200cd_dist_create", rc);

    // ---- the one thing the host language does for the library: every rank gets every rank's blob
    std::vector<uint8_t> mine(B200CD_DIST_BLOB_BYTES), all((size_t)world * B200CD_DIST_BLOB_BYTES);
    rc = b200cd_dist_export(dist, mine.data());
    if (rc != B200CD_OK) return fail("b200cd_dist_export", rc);
    if (!write_all(up, mine.data(), mine.size()) || !read_all(down, all.data(), all.size())) return fail("blob exchange", B200CD_E_IO);
    rc = b200cd_dist_connect(dist, all.data());
    if (rc != B200CD_OK) return fail("b200cd_dist_connect", rc);
    char token = 'c';  // barrier: nobody steps before everybody has mapped everybody
    if (!write_all(up, &token, 1) || !read_all(down, &token, 1)) return fail("barrier", B200CD_E_IO);

    b200cd_params p;
    b200cd_default_params(&p);  // morton.h:45,51,57
    const void* d_pairs = nullptr;
    uint64_t count = 0;
    b200cd_dist_stats st;
    memset(&st, 0, sizeof st);
    for (int k = 0; k < steps; ++k) {
        rc = b200cd_dist_step(dist, mesh, &p, &d_pairs, &count);  // main.cu:92-146, all ranks together
        if (rc != B200CD_OK) return fail("b200cd_dist_step", rc);
    }
    b200cd_dist_get_stats(dist, &st);
    if (rank == 0) {
        std::vector<uint32_t> pairs(2 * (count ? count : 1));
        rc = b200cd_copy_to_host(ctx, pairs.data(), d_pairs, count * 8);  // ordered behind rank 0's sort on the context's stream
        if (rc != B200CD_OK) return fail("b200cd_copy_to_host", rc);
        const void* d_ids = nullptr;  // makeAndPrintSet (main.cu:33-45) on the device
        uint64_t nids = 0;
        rc = b200cd_unique_triangles_device(ctx, d_pairs, count, ntris, &d_ids, &nids);
        if (rc != B200CD_OK) return fail("b200cd_unique_triangles_device", rc);
        std::vector<uint32_t> ids(nids ? nids : 1);
        rc = b200cd_copy_to_host(ctx, ids.data(), d_ids, nids * 4);
        if (rc != B200CD_OK) return fail("b200cd_copy_to_host", rc);
        printf("\nObj File Loaded:\n- %u vertexes loaded\n- %u triangles loaded\n", nverts, ntris);
        printf("%d ranks, last step %.3f ms on rank 0 (%u triangles in its Morton range)\n", world, st.ms_step, st.local_triangles);
        printf("\n\n- contact val = %llu\n", (unsigned long long)count);
        printf("\nCollision pair (%llu triangle pairs in total):\n", (unsigned long long)count);
        for (uint64_t i = 0; i < count; ++i) printf("%07u - %07u\n", pairs[2 * i], pairs[2 * i + 1]);
        printf("\n\nCollision Triangles:(%llu points in total):\n", (unsigned long long)nids);
        for (uint64_t i = 0; i < nids; ++i) printf("%u\n", ids[i]);
        printf("- Successfully Return\n");
        fflush(stdout);
    }
    token = 'd';  // barrier: peers may still be storing into my buffers before this
    if (!write_all(up, &token, 1) || !read_all(down, &token, 1)) return fail("barrier", B200CD_E_IO);
    b200cd_dist_destroy(dist);
    b200cd_mesh_destroy(mesh);
    b200cd_destroy(ctx);
    return 0;
}

int main(int argc, char** argv) {
    const char* path = nullptr;
    int world = 2, steps = 1, device = -1;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--ranks") && i + 1 < argc) world = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--steps") && i + 1 < argc) steps = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
        else path = argv[i];
    }
    if (!path || world < 1 || world > 16 || steps < 1) {
        fprintf(stderr, "usage: b200cd_dist_run mesh.obj [--ranks W (1..16)] [--steps K] [--device D]\n");
        return 2;
    }
    signal(SIGPIPE, SIG_IGN);  // a rank that died shows up as a failed write, not as a signal
    // fork BEFORE any CUDA call: every rank gets a process (and a CUDA context) of its own
    std::vector<pid_t> pid(world);
    std::vector<int> up(world), down(world);  // parent's ends: read from up[r], write to down[r]
    for (int r = 0; r < world; ++r) {
        int a[2], b[2];
        if (pipe(a) != 0 || pipe(b) != 0) { perror("pipe"); return 1; }
        pid[r] = fork();
        if (pid[r] < 0) { perror("fork"); return 1; }
        if (pid[r] == 0) {
            close(a[0]);
            close(b[1]);
            for (int q = 0; q < r; ++q) { close(up[q]); close(down[q]); }
            _exit(rank_main(r, world, path, steps, device, a[1], b[0]));
        }
        close(a[1]);
        close(b[0]);
        up[r] = a[0];
        down[r] = b[1];
    }
    bool ok = true;
    std::vector<uint8_t> all((size_t)world * B200CD_DIST_BLOB_BYTES);
    for (int r = 0; r < world && ok; ++r) ok = read_all(up[r], all.data() + (size_t)r * B200CD_DIST_BLOB_BYTES, B200CD_DIST_BLOB_BYTES);
    for (int r = 0; r < world && ok; ++r) ok = write_all(down[r], all.data(), all.size());
    for (int phase = 0; phase < 2 && ok; ++phase) {  // two barriers: "connected" and "done"
        char t;
        for (int r = 0; r < world && ok; ++r) ok = read_all(up[r], &t, 1);
        for (int r = 0; r < world && ok; ++r) ok = write_all(down[r], &t, 1);
    }
    for (int r = 0; r < world; ++r) { close(up[r]); close(down[r]); }  // a failed rank: the others see EOF and leave
    int status = 0, worst = ok ? 0 : 1;
    for (int r = 0; r < world; ++r)
        if (waitpid(pid[r], &status, 0) < 0 || !WIFEXITED(status) || WEXITSTATUS(status) != 0) worst = 1;
    return worst;
}
