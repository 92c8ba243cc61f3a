/*
 * Generate the multiply-with-carry multiplier table used by the MWC RNG streams.
 *
 * Definition (matches the table the reference ships as cuburn/code/primes.bin,
 * produced by helpers/genprimes.c:35-52): walk a = 2^32-1, 2^32-2, ... and keep
 * every a for which p = a*2^32 - 1 is prime and (p-1)/2 = a*2^31 - 1 is prime
 * (p is a "safe prime", so the MWC generator with multiplier a has the maximal
 * period (p-1)/2).  The first `count` hits are written as little-endian u32.
 *
 * This is an independent implementation: wheel/sieve pre-filter on small primes
 * followed by a deterministic 64-bit Miller-Rabin (first 12 prime bases, valid
 * for all n < 3.3e24).  OpenMP over blocks of candidates.
 *
 *   gcc -O2 -fopenmp -o gen_mwc_multipliers gen_mwc_multipliers.c
 *   ./gen_mwc_multipliers out.bin [count=262144]
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

static inline uint64_t mulmod(uint64_t a, uint64_t b, uint64_t m) {
    return (uint64_t)(((u128)a * b) % m);
}

static uint64_t powmod(uint64_t b, uint64_t e, uint64_t m) {
    uint64_t r = 1;
    b %= m;
    while (e) {
        if (e & 1) r = mulmod(r, b, m);
        b = mulmod(b, b, m);
        e >>= 1;
    }
    return r;
}

static int is_prime_u64(uint64_t n) {
    static const uint64_t bases[12] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    if (n < 2) return 0;
    for (int i = 0; i < 12; i++) {
        if (n == bases[i]) return 1;
        if (n % bases[i] == 0) return 0;
    }
    uint64_t d = n - 1;
    int s = 0;
    while (!(d & 1)) { d >>= 1; s++; }
    for (int i = 0; i < 12; i++) {
        uint64_t x = powmod(bases[i], d, n);
        if (x == 1 || x == n - 1) continue;
        int comp = 1;
        for (int r = 1; r < s; r++) {
            x = mulmod(x, x, n);
            if (x == n - 1) { comp = 0; break; }
        }
        if (comp) return 0;
    }
    return 1;
}

#define NSMALL 600
static uint32_t small_p[NSMALL];
static int n_small;

static void init_small(void) {
    n_small = 0;
    for (uint32_t c = 3; n_small < NSMALL; c += 2) {
        int ok = 1;
        for (uint32_t d = 3; d * d <= c; d += 2)
            if (c % d == 0) { ok = 0; break; }
        if (ok) small_p[n_small++] = c;
    }
}

/* Block of candidates [hi-BLK+1, hi], processed downward. */
#define BLK (1u << 16)

int main(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s out.bin [count]\n", argv[0]);
        return 2;
    }
    uint32_t want = argc > 2 ? (uint32_t)strtoul(argv[2], 0, 10) : 262144u;
    init_small();

    uint32_t *out = malloc(sizeof(uint32_t) * (want + BLK));
    uint32_t found = 0;
    uint64_t top = 4294967295ull;

    /* per-prime residues r1 = 2^32 mod q, r2 = 2^31 mod q */
    uint32_t r32[NSMALL], r31[NSMALL];
    for (int i = 0; i < n_small; i++) {
        r32[i] = (uint32_t)((1ull << 32) % small_p[i]);
        r31[i] = (uint32_t)((1ull << 31) % small_p[i]);
    }

    const int NB = 64; /* blocks per parallel batch */
    uint32_t *hits = malloc(sizeof(uint32_t) * NB * BLK);
    uint32_t nhits[64];

    while (found < want && top > 2147483648ull + (uint64_t)NB * BLK) {
#pragma omp parallel for schedule(dynamic, 1)
        for (int b = 0; b < NB; b++) {
            uint64_t hi = top - (uint64_t)b * BLK;
            uint64_t lo = hi - BLK + 1;
            uint8_t *dead = calloc(BLK, 1);
            for (int i = 0; i < n_small; i++) {
                uint32_t q = small_p[i];
                /* a*2^32 == 1 (mod q)  or  a*2^31 == 1 (mod q)  => composite */
                uint32_t inv32 = 0, inv31 = 0;
                /* modular inverse by brute force over small q (q < 5000) */
                for (uint32_t t = 1; t < q; t++) {
                    if (!inv32 && (uint64_t)t * r32[i] % q == 1) inv32 = t;
                    if (!inv31 && (uint64_t)t * r31[i] % q == 1) inv31 = t;
                    if (inv32 && inv31) break;
                }
                uint32_t base = (uint32_t)(lo % q);
                uint32_t o32 = (inv32 + q - base) % q;
                uint32_t o31 = (inv31 + q - base) % q;
                for (uint32_t k = o32; k < BLK; k += q) dead[k] = 1;
                for (uint32_t k = o31; k < BLK; k += q) dead[k] = 1;
            }
            uint32_t n = 0;
            uint32_t *h = hits + (size_t)b * BLK;
            for (int64_t k = BLK - 1; k >= 0; k--) {
                if (dead[k]) continue;
                uint64_t a = lo + (uint64_t)k;
                uint64_t p = (a << 32) - 1;
                uint64_t sg = (a << 31) - 1;
                if (is_prime_u64(sg) && is_prime_u64(p)) h[n++] = (uint32_t)a;
            }
            nhits[b] = n;
            free(dead);
        }
        for (int b = 0; b < NB && found < want; b++) {
            uint32_t *h = hits + (size_t)b * BLK;
            for (uint32_t k = 0; k < nhits[b] && found < want; k++) out[found++] = h[k];
        }
        top -= (uint64_t)NB * BLK;
    }
    if (found < want) {
        fprintf(stderr, "only found %u multipliers\n", found);
        return 1;
    }
    FILE *fp = fopen(argv[1], "wb");
    if (!fp) { perror("fopen"); return 1; }
    for (uint32_t i = 0; i < want; i++) {
        uint8_t le[4] = {out[i] & 0xff, (out[i] >> 8) & 0xff, (out[i] >> 16) & 0xff, out[i] >> 24};
        fwrite(le, 4, 1, fp);
    }
    fclose(fp);
    fprintf(stderr, "wrote %u multipliers, first=%u last=%u\n", want, out[0], out[want - 1]);
    return 0;
}
