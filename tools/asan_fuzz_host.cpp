// ASAN/UBSAN mutation fuzz of the host parser + planner (host-only code of libjpgpu): no crash, no OOB.
#include "jpgpu.h"
#include "jpgpu_host.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include <string>
using namespace jpgpu;
static std::vector<uint8_t> slurp(const char* p) { std::vector<uint8_t> v; FILE* f = fopen(p, "rb"); if (!f) return v; fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET); v.resize(n); if (fread(v.data(), 1, n, f) != (size_t)n) v.clear(); fclose(f); return v; }
int main(int argc, char** argv) {
    std::mt19937 rng(argc > 2 ? atoi(argv[2]) : 1);
    long iters = argc > 1 ? atol(argv[1]) : 1000, ok = 0, planned = 0, scans_ok = 0;
    std::vector<std::vector<uint8_t>> bases;
    for (int i = 3; i < argc; i++) { auto v = slurp(argv[i]); if (!v.empty()) bases.push_back(v); }
    if (bases.empty()) { fprintf(stderr, "no inputs\n"); return 2; }
    for (long it = 0; it < iters; it++) {
        std::vector<uint8_t> b = bases[rng() % bases.size()];
        size_t hdr = b.size();
        for (size_t i = 0; i + 1 < b.size(); i++) if (b[i] == 0xff && b[i + 1] == 0xda) { hdr = std::min(b.size(), i + 16); break; }
        int k = 1 + rng() % 3;
        for (int m = 0; m < k && b.size() > 4; m++) {
            unsigned op = rng() % 100; size_t lim = std::min(hdr, b.size()); size_t pos = 2 + rng() % (lim > 3 ? lim - 2 : 1);
            if (pos >= b.size()) continue;
            if (op < 60) { static const uint8_t pick[4] = {0, 1, 0xff, 0x11}; b[pos] = (rng() & 1) ? pick[rng() % 4] : (uint8_t)(b[pos] ^ (1u << (rng() % 8))); }
            else if (op < 75) b.erase(b.begin() + pos, b.begin() + std::min(b.size(), pos + 1 + rng() % 8));
            else if (op < 90) { size_t n = 1 + rng() % 6; for (size_t q = 0; q < n; q++) b.insert(b.begin() + pos, (uint8_t)rng()); }
            else b.resize(2 + rng() % (b.size() - 2));
        }
        for (uint32_t ext : {0u, 1u, 2u, 3u}) for (uint32_t layout : {0u, 1u, 2u}) {
            jpgpu_image_desc d; memset(&d, 0, sizeof d);
            int st = jpgpu_parse(b.data(), b.size(), ext, layout, &d);
            if (st == 0) {
                ok++;
                uint32_t mcus, bpm, nb[4];
                jpgpu_geometry(&d, &mcus, &bpm, nb);
                if ((uint64_t)d.width * d.height <= (1u << 22)) {   // keep the gather maps of REF layout small
                    uint64_t info[8];
                    if (jpgpu_plan_info(&d, 1, info) == 0) planned++;
                }
            }
            jpgpu_image_desc ds[8]; size_t n = 0;
            int s2 = jpgpu_parse_scans(b.data(), b.size(), ext | 4u, layout, ds, 8, &n);
            if (s2 == 0 && n) {
                scans_ok++;
                uint64_t info[8];
                if ((uint64_t)ds[0].frame_width * ds[0].frame_height <= (1u << 22)) jpgpu_plan_info(ds, n, info);
            }
        }
    }
    printf("iters %ld parse-ok %ld planned %ld scans-ok %ld\n", iters, ok, planned, scans_ok);
    return 0;
}
