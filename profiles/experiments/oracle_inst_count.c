// Retired CPU instructions per game step of the CPU oracle ("reported against the oracle's own minimal count"):
// perf_event_open(PERF_COUNT_HW_INSTRUCTIONS) around one single-threaded orc_playout_philox call (user space only).
//   gcc -O2 -o oracle_inst_count oracle_inst_count.c -ldl && ./oracle_inst_count ../../oracle/liboracle.so
#define _GNU_SOURCE
#include <dlfcn.h>
#include <linux/perf_event.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/ioctl.h>
#include <sys/syscall.h>
#include <unistd.h>
typedef double (*playout_fn)(int, int, uint64_t, uint32_t, uint64_t, uint64_t, int, int32_t*, uint32_t*, int32_t*, uint8_t*, int);
int main(int argc, char** argv) {
    void* lib = dlopen(argc > 1 ? argv[1] : "oracle/liboracle.so", RTLD_NOW);
    if (!lib) { printf("{\"error\": \"%s\"}\n", dlerror()); return 1; }
    playout_fn f = (playout_fn)dlsym(lib, "orc_playout_philox");
    struct perf_event_attr pe; memset(&pe, 0, sizeof pe);
    pe.type = PERF_TYPE_HARDWARE; pe.size = sizeof pe; pe.config = PERF_COUNT_HW_INSTRUCTIONS;
    pe.disabled = 1; pe.exclude_kernel = 1; pe.exclude_hv = 1; pe.inherit = 1;
    int fd = (int)syscall(SYS_perf_event_open, &pe, 0, -1, -1, 0);
    if (fd < 0) { printf("{\"error\": \"perf_event_open refused (perf_event_paranoid / seccomp): hardware instruction counters are not available in this container\"}\n"); return 0; }
    const uint64_t n = 200000;
    int32_t* pts = malloc(n * 16); uint32_t* steps = malloc(n * 4);
    printf("{");
    for (int cfg = 0; cfg < 3; ++cfg) {
        int engine = cfg == 0 ? 0 : 1, ann = cfg == 1;
        f(engine, ann, 0xD0C05EEDull, 2, 0, 1000, 1, pts, steps, NULL, NULL, 0);          // warm
        ioctl(fd, PERF_EVENT_IOC_RESET, 0); ioctl(fd, PERF_EVENT_IOC_ENABLE, 0);
        double sec = f(engine, ann, 0xD0C05EEDull, 2, 0, n, 1, pts, steps, NULL, NULL, 0);
        ioctl(fd, PERF_EVENT_IOC_DISABLE, 0);
        long long cnt = 0; if (read(fd, &cnt, sizeof cnt) != sizeof cnt) cnt = -1;
        uint64_t tot = 0; for (uint64_t i = 0; i < n; ++i) tot += steps[i];
        printf("%s\"%s\": {\"games\": %llu, \"game_steps\": %llu, \"cpu_instructions\": %lld, \"instructions_per_game_step\": %.1f, \"sec\": %.3f}",
               cfg ? ", " : "", cfg == 0 ? "rs_doko" : (cfg == 1 ? "rs_full_doko_with_announcements" : "rs_full_doko_no_announcement_policy"),
               (unsigned long long)n, (unsigned long long)tot, cnt, (double)cnt / (double)tot, sec);
    }
    printf("}\n");
    return 0;
}
