// Dev check: the LA table builder (csrc/host/fs_host.cpp, pipelined form) under ThreadSanitizer: 20 builds of one view, every
// table must hash the same and TSAN must stay silent.  Build and run (view strings: fractalshark_b200/views.py):
//   g++ -fsanitize=thread -O1 -g -std=c++17 -ffp-contract=off -I/usr/local/cuda/include -o /tmp/drv_tsan tools/tsan_la_build.cpp \
//       fractalshark_b200/csrc/host/fs_host.cpp -l:libgmp.so.10 -pthread
//   FS_HOST_THREADS=3 /tmp/drv_tsan <minX> <minY> <maxX> <maxY> <iterations> <numeric enum>
// Round 2: views 5 with 2, 3, 8 and 16 threads -- no report, identical tables.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
struct fsh_view; struct fsh_orbit; struct fsh_la;
extern "C" {
fsh_view *fsh_view_create(const char *, const char *, const char *, const char *, uint32_t, uint32_t, uint32_t, int32_t);
fsh_orbit *fsh_orbit_compute(const fsh_view *, int32_t, uint64_t, int32_t);
fsh_la *fsh_la_build(const fsh_orbit *, uint32_t);
void fsh_la_destroy(fsh_la *);
const void *fsh_la_las(const fsh_la *);
uint64_t fsh_la_num_las(const fsh_la *);
}
int main(int argc, char **argv) {
    fsh_view *v = fsh_view_create(argv[1], argv[2], argv[3], argv[4], 96, 54, 1, 0);
    fsh_orbit *o = fsh_orbit_compute(v, atoi(argv[6]), strtoull(argv[5], 0, 10), 1);
    if (!o) { printf("no orbit\n"); return 1; }
    unsigned long long h0 = 0;
    for (int rep = 0; rep < 20; rep++) {
        fsh_la *l = fsh_la_build(o, 4);
        const unsigned char *p = (const unsigned char *)fsh_la_las(l);
        unsigned long long h = 1469598103934665603ull;
        const uint64_t n = fsh_la_num_las(l) * 68; // HDRx32 / u32 records
        for (uint64_t i = 0; i < n; i++) h = (h ^ p[i]) * 1099511628211ull;
        if (rep == 0) h0 = h; else if (h != h0) { printf("MISMATCH at rep %d\n", rep); return 2; }
        if (rep == 0) printf("records %llu hash %llx\n", (unsigned long long)fsh_la_num_las(l), h);
        fsh_la_destroy(l);
    }
    printf("ok\n");
    return 0;
}
