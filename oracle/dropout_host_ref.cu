// Host-side compilation of the product's own dropout_keep() (csrc/b2q_internal.h) -- the same source the device runs.
// Prints keep bits for a grid of (seed, p, index) so that oracle/dropout.py can be pinned against it on the CPU.
#include <cstdio>
#include <cstdlib>
#include "../causal-unified-language-vision_b200/csrc/b2q_internal.h"
// "bits" mode: checks the packed-mask helpers (dropout_bits32, dropout_byte_to_masks) against dropout_keep on the host.
static int check_bits(unsigned long long seed, float p, long long n) {
    const uint32_t thr = b2q::dropout_threshold(p);
    for (long long e0 = 0; e0 + 32 <= n; e0 += 32) {
        const uint32_t bits = b2q::dropout_bits32(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32),
                                                  static_cast<unsigned long long>(e0), thr);
        for (int i = 0; i < 32; ++i)
            if (((bits >> i) & 1u) != (b2q::dropout_keep(seed, static_cast<unsigned long long>(e0 + i), thr) ? 1u : 0u)) return 1;
    }
    for (uint32_t byte = 0; byte < 256; ++byte) {
        uint32_t m[4];
        b2q::dropout_byte_to_masks(byte, m);
        for (int j = 0; j < 4; ++j) {
            const uint32_t want = (((byte >> (2 * j)) & 1u) ? 0x0000FFFFu : 0u) | (((byte >> (2 * j + 1)) & 1u) ? 0xFFFF0000u : 0u);
            if (m[j] != want) return 2;
        }
        // the device path reads byte SIGNS of the two products: check them too (what PRMT's sign-replicate mode sees)
        const uint32_t x0 = (byte & 0xFu) * 0x10204080u, x1 = ((byte >> 4) & 0xFu) * 0x10204080u;
        for (int k = 0; k < 4; ++k) {
            if (((x0 >> (8 * k + 7)) & 1u) != ((byte >> k) & 1u)) return 3;
            if (((x1 >> (8 * k + 7)) & 1u) != ((byte >> (4 + k)) & 1u)) return 3;
        }
    }
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 4 && argv[4][0] == 'b') {
        const int rc = check_bits(strtoull(argv[1], nullptr, 0), static_cast<float>(atof(argv[2])), atoll(argv[3]));
        printf(rc == 0 ? "bits ok\n" : "bits MISMATCH %d\n", rc);
        return rc;
    }
    const unsigned long long seed = strtoull(argv[1], nullptr, 0);
    const float p = static_cast<float>(atof(argv[2]));
    const long long n = atoll(argv[3]);
    const uint32_t thr = b2q::dropout_threshold(p);
    for (long long i = 0; i < n; ++i) putchar(b2q::dropout_keep(seed, static_cast<unsigned long long>(i), thr) ? '1' : '0');
    putchar('\n');
    return 0;
}
