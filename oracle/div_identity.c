/* Test infrastructure (not product code): checks on the host the identity the mid-stage kernels
 * rely on (warpstr_b200/csrc/wstr_internal.h, div_fast: the mid-stage kernels and the normalisation kernel): with y = RN(1/d),
 *     q0 = RN(a*y); r0 = fma(-q0, d, a); q1 = fma(r0, y, q0); r1 = fma(-q1, d, a); q = fma(r1, y, q1)
 * equals the IEEE quotient a/d bit for bit (Markstein's correction step applied twice).  Operands: random
 * significands plus adversarial ones (divisor all ones / a power of two / 1.5 = the division by 3),
 * exponents inside the range the kernels allow; and the normalisation kernel's (sample - shift) / scale.  Prints the mismatch counts of the one- and two-step forms.
 * Build: gcc -O2 -mfma -ffp-contract=off -o div_identity div_identity.c -lm ; run: ./div_identity [pairs] */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t s[4] = {0x9E3779B97F4A7C15ull, 0xBF58476D1CE4E5B9ull, 0x94D049BB133111EBull, 0x2545F4914F6CDD1Dull};
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t next(void) {   /* xoshiro256** */
    const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
}
static inline double make(uint64_t mant, int e, int neg) {
    const uint64_t b = ((uint64_t)neg << 63) | ((uint64_t)(1023 + e) << 52) | (mant & 0xfffffffffffffull);
    double d;
    memcpy(&d, &b, 8);
    return d;
}

int main(int argc, char **argv) {
    const long n = argc > 1 ? atol(argv[1]) : 20000000L;
    long bad1 = 0, bad2 = 0;
    for (long i = 0; i < n; ++i) {
        uint64_t ma = next(), md = next();
        const uint64_t ee = next();
        const int mode = (int)(ee & 7);
        if (mode == 1) md |= 0xffffffffff000ull;            /* divisor significand close to all ones */
        if (mode == 2) md &= 0xfffull;                      /* divisor close to a power of two */
        if (mode == 3) ma |= 0xfffffffff0000ull;
        if (mode == 4) { ma &= 0xffffull; md &= 0xffffull; }
        if (mode == 5) md = 0x8000000000000ull;             /* 1.5 * 2^e: the divisions by 3 */
        const int ea = (int)((ee >> 8) % 1201) - 600, ed = (int)((ee >> 24) % 121) - 60;
        volatile double a = make(ma, ea, (int)((ee >> 40) & 1)), d = make(md, ed, (int)((ee >> 41) & 1));
        if (mode == 6 || mode == 7) {
            /* the normalisation kernel's operands (aux.cu): (x - shift) / scale, x an int16 sample, shift a
             * mean of two interpolated percentiles, scale a median absolute deviation: an integer, a half, or
             * (mode 7) anything of that size */
            const double shift = (double)(int)(ma % 2048) + (mode == 6 ? (double)((ma >> 16) % 4) * 0.25 : make(ma >> 12, -1, 0) - 0.5);
            const double scale = mode == 6 ? (double)(1 + (int)(md % 400)) * 0.5 : make(md, (int)(md % 9), 0);
            a = (double)(int)((ee >> 44) % 65536 - 32768) - shift;
            d = scale;
        }
        const double y = 1.0 / d, ref = a / d;
        const double q0 = a * y;
        const double r0 = fma(-q0, d, a);
        const double q1 = fma(r0, y, q0);
        const double r1 = fma(-q1, d, a);
        const double q2 = fma(r1, y, q1);
        if (q1 != ref) ++bad1;
        if (q2 != ref) {
            if (bad2 < 5) printf("mismatch a=%a d=%a got=%a want=%a\n", a, d, q2, ref);
            ++bad2;
        }
    }
    printf("pairs=%ld one_step_mismatches=%ld two_step_mismatches=%ld\n", n, bad1, bad2);
    return bad2 != 0;
}
