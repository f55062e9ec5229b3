// Host-side compilation of the product's own dropout_keep() (csrc/b2q_internal.h) -- the same source the device runs.
// Prints keep bits for a grid of (seed, p, index) so that oracle/dropout.py can be pinned against it on the CPU.
#include <cstdio>
#include <cstdlib>
#include "../causal-unified-language-vision_b200/csrc/b2q_internal.h"
int main(int argc, char** argv) {
    const unsigned long long seed = strtoull(argv[1], nullptr, 0);
    const float p = static_cast<float>(atof(argv[2]));
    const long long n = atoll(argv[3]);
    const uint32_t thr = b2q::dropout_threshold(p);
    for (long long i = 0; i < n; ++i) putchar(b2q::dropout_keep(seed, static_cast<unsigned long long>(i), thr) ? '1' : '0');
    putchar('\n');
    return 0;
}
