// CPU check of tools/experiments/float_advance.h against plain repeated float32 addition.
// Build: g++ -O2 -ffp-contract=off -msse2 -mfpmath=sse.  Prints "ok <cases>" or the first mismatch.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <initializer_list>
#include "float_advance.h"

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 16); }
static float seq(float p, float s, int k) { volatile float a = p; for (int i = 0; i < k; ++i) a = a + s; return a; }
static bool same(float a, float b) { return tsdf::float_bits(a) == tsdf::float_bits(b) || (a != a && b != b); }

static long long cases = 0;
static int check1(float p, float s, int k) {
  ++cases;
  const float want = seq(p, s, k), got = tsdf::advance_exact(p, s, k);
  if (!same(want, got)) { printf("MISMATCH p=%a s=%a k=%d want=%a got=%a\n", p, s, k, want, got); return 1; }
  return 0;
}

int main(int argc, char** argv) {
  const long long n_random = argc > 1 ? atoll(argv[1]) : 2000000;
  int bad = 0;
  // 1. random magnitudes typical of voxel coordinates and steps, all signs
  for (long long i = 0; i < n_random && !bad; ++i) {
    const int ep = (int)(rnd() % 20) - 4, es = (int)(rnd() % 12) - 8;
    float p = ldexpf(1.0f + (rnd() & 0x7FFFFF) / 8388608.0f, ep), s = ldexpf(1.0f + (rnd() & 0x7FFFFF) / 8388608.0f, es);
    if (rnd() & 1) p = -p;
    if (rnd() & 1) s = -s;
    bad |= check1(p, s, (int)(rnd() % 700));
  }
  // 2. ties at every step: s = (S + 1/2) ulp(p) for every small S, odd and even starting mantissas, both directions
  for (int e = -3; e < 16 && !bad; ++e)
    for (int S = 0; S < 40 && !bad; ++S)
      for (int m0 = 0; m0 < 8 && !bad; ++m0)
        for (int sg = 0; sg < 4 && !bad; ++sg) {
          const float ulp = ldexpf(1.0f, e - 23);
          float p = ldexpf(1.0f, e) + (float)(m0 + ((rnd() % 1000) << 3)) * ulp, s = ((float)S + 0.5f) * ulp;
          if (sg & 1) p = -p;
          if (sg & 2) s = -s;
          for (int k : {0, 1, 2, 3, 5, 6, 7, 8, 9, 33, 100, 517}) bad |= check1(p, s, k);
        }
  // 3. binade edges: start a few ulps below / above a power of two and walk across it, fine and coarse steps
  for (int e = -2; e < 16 && !bad; ++e)
    for (int off = -6; off <= 6 && !bad; ++off)
      for (int q = 1; q < 64 && !bad; ++q) {
        const float ulp = ldexpf(1.0f, e - 24);
        const float p = ldexpf(1.0f, e) + (float)off * ulp;
        for (float s : {(float)q * 0.25f * ulp, (float)q * 0.37f * ulp, -(float)q * 0.25f * ulp, -(float)q * 0.61f * ulp, (float)q * 1.5f * ulp})
          for (int k : {7, 8, 15, 64, 300}) { bad |= check1(p, s, k); bad |= check1(-p, s, k); }
      }
  // 4. through zero, stuck accumulators, huge steps, specials
  for (int i = 0; i < 200000 && !bad; ++i) {
    const float p = ((float)(rnd() % 2001) - 1000.0f) * 0.01f, s = ((float)(rnd() % 2001) - 1000.0f) * ldexpf(1.0f, -(int)(rnd() % 30));
    bad |= check1(p, s, (int)(rnd() % 400));
  }
  bad |= check1(0.0f, 0.0f, 100); bad |= check1(-0.0f, 0.0f, 100); bad |= check1(1e30f, 1e38f, 50); bad |= check1(1.0f, 1e-30f, 500);
  bad |= check1(3.0e38f, 3.0e38f, 9); bad |= check1(1.0f, NAN, 20); bad |= check1(1e-40f, 1e-42f, 300); bad |= check1(1e-38f, -1e-40f, 300);
  // 5. the three-accumulator form
  for (int i = 0; i < 400000 && !bad; ++i) {
    float p[3], s[3];
    for (int c = 0; c < 3; ++c) {
      p[c] = ((float)(rnd() % 200001) - 100000.0f) * 0.01f;
      s[c] = ((float)(rnd() % 60001) - 30000.0f) * 1e-4f;
      if ((rnd() & 15) == 0) s[c] = 0.0f;
    }
    const int k = (int)(rnd() % 600);
    float x = p[0], y = p[1], z = p[2];
    tsdf::advance_exact3(x, y, z, s[0], s[1], s[2], k);
    ++cases;
    if (!same(x, seq(p[0], s[0], k)) || !same(y, seq(p[1], s[1], k)) || !same(z, seq(p[2], s[2], k))) {
      printf("MISMATCH3 p=(%a,%a,%a) s=(%a,%a,%a) k=%d\n", p[0], p[1], p[2], s[0], s[1], s[2], k); bad = 1;
    }
  }
  if (!bad) printf("ok %lld\n", cases);
  return bad;
}
