// Sanitizer fuzz of the sequence reader's zip handling (csrc/mdc_sequence.cpp: end-of-central-directory search, central directory,
// local headers, stored and deflated entries): a small valid images.zip is damaged in seeded random ways, opened with mdc_seq_open and
// every entry read back.  Only survival under -fsanitize=address,undefined is checked.  Built and run by tests/test_decoder_fuzz.py.
// usage: zip_fuzz <sequence dir with images.zip + times.txt> <rounds>     (the directory's images.zip is overwritten)
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <random>
#include <string>
#include <vector>

#include "mdc_b200.h"
#include "mdc_internal.h"

// the GPU side of the library is not linked here
void mdc_set_error(const char*, ...) {}
void mdc_ctx_geometry(const mdc_ctx*, int* a, int* b, int* c, int* d) { *a = *b = *c = *d = 0; }
extern "C" const char* mdc_last_error(void) { return ""; }
extern "C" int mdc_host_alloc(void** p, size_t n) { *p = malloc(n); return *p ? 0 : 1; }
extern "C" void mdc_host_free(void* p) { free(p); }
extern "C" int mdc_prepare_batch_host(mdc_ctx*, const uint8_t*, int, unsigned, float* const*, int) { return 1; }

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: zip_fuzz <dir> <rounds>\n"); return 2; }
    const std::string dir = argv[1], zip = dir + "/images.zip";
    std::vector<uint8_t> orig;
    {
        std::ifstream f(zip, std::ios::binary);
        orig.assign((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    }
    if (orig.size() < 64) return 2;
    const int rounds = atoi(argv[2]);
    std::mt19937 rng(777);
    std::vector<uint8_t> px(1 << 20);
    int opened = 0, frames = 0;
    for (int r = 0; r <= rounds; ++r) {
        std::vector<uint8_t> b = orig;
        if (r > 0) switch (r % 4) {
            case 0: for (int k = 0; k < 1 + static_cast<int>(rng() % 8); ++k) b[rng() % b.size()] = static_cast<uint8_t>(rng()); break;
            case 1: for (int k = 0; k < 1 + static_cast<int>(rng() % 6); ++k) b[b.size() - 1 - rng() % std::min<size_t>(b.size(), 400)] = static_cast<uint8_t>(rng()); break;   // directory area
            case 2: b.resize(rng() % b.size()); break;
            default: {
                const size_t p = rng() % b.size(), n = std::min<size_t>(b.size() - p, 1 + rng() % 32);
                for (size_t i = 0; i < n; ++i) b[p + i] = 0xff;
            }
        }
        {
            std::ofstream o(zip, std::ios::binary | std::ios::trunc);
            o.write(reinterpret_cast<const char*>(b.data()), static_cast<std::streamsize>(b.size()));
        }
        mdc_seq* s = nullptr;
        if (mdc_seq_open(dir.c_str(), &s) != 0 || !s) { if (r == 0) { fprintf(stderr, "the intact archive does not open\n"); return 3; } continue; }
        ++opened;
        const int n = mdc_seq_num_images(s);
        for (int i = 0; i < n && i < 16; ++i) {
            int w = 0, h = 0;
            if (mdc_seq_read_gray8(s, i, px.data(), px.size(), &w, &h) == 0) ++frames;
        }
        mdc_seq_close(s);
    }
    {   // leave the directory as it was
        std::ofstream o(zip, std::ios::binary | std::ios::trunc);
        o.write(reinterpret_cast<const char*>(orig.data()), static_cast<std::streamsize>(orig.size()));
    }
    printf("%d rounds, %d archives opened, %d frames read\n", rounds, opened, frames);
    return 0;
}
