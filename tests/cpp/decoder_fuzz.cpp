// Sanitizer fuzz of the host image decoders (csrc/mdc_jpeg.cpp, csrc/mdc_gray_image.cpp): a valid file is damaged in seeded random ways
// (byte flips anywhere, in the header / table area, truncation, runs of 0xFF) and decoded; the only thing checked is that the process
// survives under -fsanitize=address,undefined.  Built and run by tests/test_decoder_fuzz.py.  usage: decoder_fuzz <file> <rounds>
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <random>
#include <string>
#include <vector>

#include "mdc_internal.h"

void mdc_set_error(const char*, ...) {}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: decoder_fuzz <file> <rounds>\n"); return 2; }
    std::ifstream f(argv[1], std::ios::binary);
    const std::vector<uint8_t> orig((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (orig.size() < 16) return 2;
    const int rounds = atoi(argv[2]);
    std::mt19937 rng(12345);
    mdc_gray_image img;
    int ok = mdc_decode_gray_image(orig, "intact", &img) ? 1 : 0;
    if (!ok) { fprintf(stderr, "the intact file does not decode\n"); return 3; }
    for (int r = 0; r < rounds; ++r) {
        std::vector<uint8_t> b = orig;
        switch (r % 4) {
            case 0: for (int k = 0; k < 1 + static_cast<int>(rng() % 8); ++k) b[rng() % b.size()] = static_cast<uint8_t>(rng()); break;
            case 1: for (int k = 0; k < 1 + static_cast<int>(rng() % 4); ++k) b[rng() % std::min<size_t>(b.size(), 700)] = static_cast<uint8_t>(rng()); break;
            case 2: b.resize(rng() % b.size()); break;
            default: {
                const size_t p = rng() % b.size(), n = std::min<size_t>(b.size() - p, 1 + rng() % 64);
                for (size_t i = 0; i < n; ++i) b[p + i] = 0xff;
            }
        }
        ok += mdc_decode_gray_image(b, "fuzz", &img) ? 1 : 0;
    }
    printf("%d rounds, %d decoded\n", rounds, ok);
    return 0;
}
