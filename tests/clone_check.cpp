// Host-side check of groundsegmentationrosnode_b200/csrc/gs_math.cuh (compiled with GS_MATH_HOST) against the
// box's libm: gs_atan2f must be bit-identical to atan2f.  Prints "n=<count> bad=<mismatches>".
#define GS_MATH_HOST 1
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "../groundsegmentationrosnode_b200/csrc/gs_math.cuh"

int main(int argc, char** argv) {
    const long total = argc > 1 ? atol(argv[1]) : 10000000L;
    uint64_t s = 88172645463325252ULL;
    long bad = 0, n = 0;
    for (long i = 0; i < total; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        float y, x;
        switch (i & 3) {
            case 0:  // the angle image's domain: non-negative metre-scale deltas
                y = (float)((s & 0xffffff) / 16777216.0 * 60.0);
                x = (float)(((s >> 24) & 0xffffff) / 16777216.0 * 60.0);
                break;
            case 1:  // tiny over large and large over tiny
                y = (float)((s & 0xffffff) / 16777216.0 * 1e-3);
                x = (float)(((s >> 24) & 0xffffff) / 16777216.0 * 100.0);
                if (i & 4) { float t = x; x = y; y = t; }
                break;
            case 2:  // all signs
                y = (float)(((s & 0xffffff) / 16777216.0 - 0.5) * 200.0);
                x = (float)((((s >> 24) & 0xffffff) / 16777216.0 - 0.5) * 200.0);
                break;
            default:  // arbitrary bit patterns (incl. zeros, denormals, infinities; NaNs skipped)
                y = gs_i2f((int32_t)(uint32_t)s);
                x = gs_i2f((int32_t)(uint32_t)(s >> 32));
                if (y != y || x != x) continue;
        }
        const float a = atan2f(y, x), b = gs_atan2f(y, x);
        ++n;
        if (gs_f2i(a) != gs_f2i(b)) {
            if (bad < 5) printf("y=%a x=%a libm=%a clone=%a\n", y, x, a, b);
            ++bad;
        }
    }
    const float sp[][2] = {{0.f, 0.f}, {0.f, 1.f}, {1.f, 0.f}, {0.f, -1.f}, {-0.f, -1.f}, {1.f, 1.f}, {2.f, 1.f}, {1e-20f, 1.f}, {5.f, 1e-30f}};
    for (unsigned i = 0; i < sizeof(sp) / sizeof(sp[0]); ++i) {
        ++n;
        if (gs_f2i(atan2f(sp[i][0], sp[i][1])) != gs_f2i(gs_atan2f(sp[i][0], sp[i][1]))) ++bad;
    }
    // decode_range (gs_fused.cu): n / 65534.0 as q0 = n*R, r = fma(-q0, 65534, n), q = fma(r, R, q0) — exhaustive
    long bad_div = 0;
    {
        const double R = 1.0 / 65534.0;
        for (int k = 0; k <= 65534; ++k) {
            const double nn = (double)k;
            const double q0 = nn * R;
            const double r = fma(-q0, 65534.0, nn);
            const double q = fma(r, R, q0);
            if (q != nn / 65534.0) ++bad_div;
        }
    }
    printf("n=%ld bad=%ld bad_div=%ld\n", n, bad, bad_div);
    return (bad || bad_div) ? 1 : 0;
}
