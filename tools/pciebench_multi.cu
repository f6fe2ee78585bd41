// tools/pciebench_multi.cu -- aggregate host<->device copy ceiling of the box over 1..N GPUs.
//
// The end-to-end number of bench.py (fmd_demod_host: pinned host in -> H2D -> kernel -> D2H -> pinned
// host out) is bounded by what the host side of the box can move, not by the GPUs.  This tool
// measures that bound with copies only (no kernels): N workers, one per GPU, each moving its own
// pinned buffers H2D, D2H or both at once, started together, timed by the wall clock from the
// common start to the last worker's finish.
//
//   pciebench_multi [--workers threads|procs] [--alloc pinned|wc|thp|hugetlb] [--mib 512] [--slice 32]
//                   [--reps 8] [--counts 1,2,4,8]
//
//   threads : one process, one thread per GPU (one CUDA context per device, shared address space)
//   procs   : one forked process per GPU (what torchrun does), synchronised through shared memory
//   alloc   : pinned  = cudaHostAlloc(Portable)
//             wc      = H2D source write-combined (cudaHostAllocWriteCombined), D2H target plain
//             thp     = mmap + MADV_HUGEPAGE + first touch + cudaHostRegister (2 MiB pages: fewer IOMMU entries)
//             hugetlb = mmap(MAP_HUGETLB) + cudaHostRegister (needs reserved huge pages; reported if refused)
//
// Output: one line per (count, direction): aggregate GB/s per direction.
#include <cuda_runtime.h>
#include <pthread.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));  \
            exit(3);                                                                            \
        }                                                                                       \
    } while (0)

static double now() {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double) t.tv_sec + 1e-9 * (double) t.tv_nsec;
}

struct Shared {
    pthread_barrier_t bar;
    double t_start[16], t_end[16];
};

struct Opts {
    std::string alloc = "pinned";
    size_t bytes = 512ull << 20, slice = 32ull << 20;
    int reps = 8;
};

static void *host_alloc(const Opts &o, size_t n, bool h2d_source) {
    void *p = nullptr;
    if (o.alloc == "pinned" || (o.alloc == "wc" && !h2d_source)) {
        CK(cudaHostAlloc(&p, n, cudaHostAllocPortable));
    } else if (o.alloc == "wc") {
        CK(cudaHostAlloc(&p, n, cudaHostAllocPortable | cudaHostAllocWriteCombined));
    } else {
        int flags = MAP_PRIVATE | MAP_ANONYMOUS;
        if (o.alloc == "hugetlb") flags |= MAP_HUGETLB;
        p = mmap(nullptr, n, PROT_READ | PROT_WRITE, flags, -1, 0);
        if (p == MAP_FAILED) {
            perror("mmap");
            exit(4);
        }
        if (o.alloc == "thp") madvise(p, n, MADV_HUGEPAGE);
        memset(p, 1, n); /* first touch before pinning, so the huge pages exist */
        CK(cudaHostRegister(p, n, cudaHostRegisterPortable));
        return p;
    }
    memset(p, 1, n);
    return p;
}

/* one worker: device `dev`, directions: bit0 = H2D, bit1 = D2H */
static void worker(int idx, int dev, int dirs_list_n, const int *dirs_list, const Opts &o, Shared *sh) {
    CK(cudaSetDevice(dev));
    void *h_in = host_alloc(o, o.bytes, true), *h_out = host_alloc(o, o.bytes, false), *d_in, *d_out;
    CK(cudaMalloc(&d_in, o.bytes));
    CK(cudaMalloc(&d_out, o.bytes));
    CK(cudaMemset(d_out, 2, o.bytes));
    cudaStream_t a, b;
    CK(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking));
    for (int k = 0; k < dirs_list_n; ++k) {
        const int dirs = dirs_list[k];
        for (int pass = 0; pass < 2; ++pass) { /* pass 0 warms up */
            CK(cudaDeviceSynchronize());
            pthread_barrier_wait(&sh->bar);
            const double t0 = now();
            const int reps = pass ? o.reps : 1;
            for (int r = 0; r < reps; ++r)
                for (size_t off = 0; off < o.bytes; off += o.slice) {
                    const size_t len = o.bytes - off < o.slice ? o.bytes - off : o.slice;
                    if (dirs & 1) CK(cudaMemcpyAsync((char *) d_in + off, (char *) h_in + off, len, cudaMemcpyHostToDevice, a));
                    if (dirs & 2) CK(cudaMemcpyAsync((char *) h_out + off, (char *) d_out + off, len, cudaMemcpyDeviceToHost, b));
                }
            CK(cudaStreamSynchronize(a));
            CK(cudaStreamSynchronize(b));
            const double t1 = now();
            if (pass) {
                sh->t_start[idx] = t0;
                sh->t_end[idx] = t1;
            }
            pthread_barrier_wait(&sh->bar);
        }
        pthread_barrier_wait(&sh->bar); /* the reporter reads the times between these two */
    }
    if (o.alloc == "thp" || o.alloc == "hugetlb") {
        cudaHostUnregister(h_in);
        cudaHostUnregister(h_out);
        munmap(h_in, o.bytes);
        munmap(h_out, o.bytes);
    } else {
        cudaFreeHost(h_in);
        cudaFreeHost(h_out);
    }
    cudaFree(d_in);
    cudaFree(d_out);
}

struct ThreadArg {
    int idx, dev, nd;
    const int *dirs;
    const Opts *o;
    Shared *sh;
};
static void *thread_main(void *p) {
    ThreadArg *a = (ThreadArg *) p;
    worker(a->idx, a->dev, a->nd, a->dirs, *a->o, a->sh);
    return nullptr;
}

/* the reporter takes part in every barrier and prints after each measured pass */
static void report(int n, int nd, const int *dirs, const Opts &o, Shared *sh, const char *workers) {
    static const char *names[4] = {"", "H2D only", "D2H only", "H2D + D2H"};
    for (int k = 0; k < nd; ++k) {
        for (int pass = 0; pass < 2; ++pass) {
            pthread_barrier_wait(&sh->bar);
            pthread_barrier_wait(&sh->bar);
        }
        double t0 = 1e300, t1 = 0, slowest = 0, fastest = 1e300;
        for (int i = 0; i < n; ++i) {
            if (sh->t_start[i] < t0) t0 = sh->t_start[i];
            if (sh->t_end[i] > t1) t1 = sh->t_end[i];
            const double d = sh->t_end[i] - sh->t_start[i];
            if (d > slowest) slowest = d;
            if (d < fastest) fastest = d;
        }
        const double per_dir = (double) n * (double) o.bytes * o.reps / (t1 - t0) / 1e9;
        printf("%-7s alloc=%-7s slice=%3zuMiB gpus=%d %-10s %7.1f GB/s per direction aggregate (%6.1f per GPU; "
               "worker spread %.0f..%.0f ms)\n",
               workers, o.alloc.c_str(), o.slice >> 20, n, names[dirs[k]], per_dir, per_dir / n, fastest * 1e3,
               slowest * 1e3);
        fflush(stdout);
        pthread_barrier_wait(&sh->bar);
    }
}

int main(int argc, char **argv) {
    Opts o;
    std::string workers = "threads", counts = "1,2,4,8";
    for (int i = 1; i + 1 < argc; i += 2) {
        const std::string k = argv[i], v = argv[i + 1];
        if (k == "--workers") workers = v;
        else if (k == "--alloc") o.alloc = v;
        else if (k == "--mib") o.bytes = (size_t) atol(v.c_str()) << 20;
        else if (k == "--slice") o.slice = (size_t) atol(v.c_str()) << 20;
        else if (k == "--reps") o.reps = atoi(v.c_str());
        else if (k == "--counts") counts = v;
        else {
            fprintf(stderr, "unknown option %s\n", k.c_str());
            return 2;
        }
    }
    const int dirs[3] = {1, 2, 3};
    std::vector<int> ns;
    for (char *s = strdup(counts.c_str()), *t = strtok(s, ","); t; t = strtok(nullptr, ",")) ns.push_back(atoi(t));

    /* the device count must be known without creating a context in the parent of forked workers */
    int ndev = 0;
    if (workers == "procs") {
        FILE *p = popen("nvidia-smi -L 2>/dev/null | wc -l", "r");
        if (p) {
            if (fscanf(p, "%d", &ndev) != 1) ndev = 0;
            pclose(p);
        }
    } else {
        CK(cudaGetDeviceCount(&ndev));
    }
    if (ndev <= 0) {
        fprintf(stderr, "no GPU\n");
        return 1;
    }
    for (int n : ns) {
        if (n > ndev || n > 16) continue;
        Shared *sh = (Shared *) mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
        pthread_barrierattr_t at;
        pthread_barrierattr_init(&at);
        pthread_barrierattr_setpshared(&at, PTHREAD_PROCESS_SHARED);
        pthread_barrier_init(&sh->bar, &at, (unsigned) n + 1);
        if (workers == "procs") {
            std::vector<pid_t> kids;
            for (int i = 0; i < n; ++i) {
                pid_t pid = fork();
                if (pid == 0) {
                    worker(i, i, 3, dirs, o, sh);
                    _exit(0);
                }
                kids.push_back(pid);
            }
            report(n, 3, dirs, o, sh, "procs");
            for (pid_t k : kids) waitpid(k, nullptr, 0);
        } else {
            std::vector<pthread_t> th(n);
            std::vector<ThreadArg> ta(n);
            for (int i = 0; i < n; ++i) {
                ta[i] = {i, i, 3, dirs, &o, sh};
                pthread_create(&th[i], nullptr, thread_main, &ta[i]);
            }
            report(n, 3, dirs, o, sh, "threads");
            for (int i = 0; i < n; ++i) pthread_join(th[i], nullptr);
        }
        munmap(sh, sizeof(Shared));
    }
    return 0;
}
