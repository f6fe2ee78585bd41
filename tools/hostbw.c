/* tools/hostbw.c -- what the host's memory system moves: T threads, each memcpy-ing (or just
 * reading / just writing) its own 256 MiB buffers.  Context for the end-to-end ceiling of the
 * box: every byte that crosses PCIe is also read from or written to host DRAM.
 *   hostbw [threads=1,2,4,8,16,32]      prints GB/s of payload (copy moves 2x that over the bus) */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define N (256ull << 20)
static pthread_barrier_t bar;
static int mode; /* 0 copy, 1 read, 2 write */
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
static volatile uint64_t sink;

static void *work(void *p) {
    (void) p;
    char *a = malloc(N), *b = malloc(N);
    memset(a, 1, N); memset(b, 2, N);
    for (int m = 0; m < 3; ++m) {
        pthread_barrier_wait(&bar);
        for (int r = 0; r < 4; ++r) {
            if (m == 0) memcpy(b, a, N);
            else if (m == 1) { uint64_t s = 0; const uint64_t *q = (const uint64_t *) a; for (size_t i = 0; i < N / 8; i += 1) s += q[i]; sink += s; }
            else memset(b, r, N);
        }
        pthread_barrier_wait(&bar);
    }
    free(a); free(b);
    return NULL;
}

int main(int argc, char **argv) {
    const char *list = argc > 1 ? argv[1] : "1,2,4,8,16,32";
    char *dup = strdup(list);
    for (char *t = strtok(dup, ","); t; t = strtok(NULL, ",")) {
        int T = atoi(t);
        if (T < 1 || T > 256) continue;
        pthread_t th[256];
        pthread_barrier_init(&bar, NULL, (unsigned) T + 1);
        for (int i = 0; i < T; ++i) pthread_create(&th[i], NULL, work, NULL);
        static const char *nm[3] = {"memcpy", "read", "memset"};
        for (mode = 0; mode < 3; ++mode) {
            pthread_barrier_wait(&bar);
            double t0 = now();
            pthread_barrier_wait(&bar);
            double t1 = now();
            printf("host %-6s threads=%-3d %7.1f GB/s payload\n", nm[mode], T, 4.0 * T * (double) N / (t1 - t0) / 1e9);
        }
        for (int i = 0; i < T; ++i) pthread_join(th[i], NULL);
        pthread_barrier_destroy(&bar);
        fflush(stdout);
    }
    return 0;
}
