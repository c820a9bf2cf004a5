// Host check: the integer fast paths of evg_score.cuh equal the literal FP64 formulas.
#include <cstdio>
#include <initializer_list>
#include <cstdlib>
#include <cmath>
#include "evergreen_b200/csrc/evg_score.cuh"
using namespace evg;
static int64_t ref_floor_minutes_over(int64_t d, int64_t n) { return int64_t(std::floor((double(d / kMinute) + double(d % kMinute) / (60.0 * 1e9)) / double(n))); }
static int64_t ref_trunc_hours(int64_t d) { return int64_t(double(d / kHour) + double(d % kHour) / (3600.0 * 1e9)); }
int main() {
  uint64_t x = 88172645463325252ULL; long bad = 0; long n = 0;
  auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
  for (int it = 0; it < 20000000; it++) {
    int64_t q = int64_t(rnd() % (uint64_t(1) << 15));
    int64_t r; switch (rnd() % 4) { case 0: r = kMinute - 1 - int64_t(rnd() % 3); break; case 1: r = int64_t(rnd() % 3); break; default: r = int64_t(rnd() % uint64_t(kMinute)); }
    int64_t d = q * kMinute + r;
    if (floor_minutes_over(d, 1) != ref_floor_minutes_over(d, 1)) bad++;
    int64_t rh; switch (rnd() % 4) { case 0: rh = kHour - 1 - int64_t(rnd() % 3); break; case 1: rh = int64_t(rnd() % 3); break; default: rh = int64_t(rnd() % uint64_t(kHour)); }
    int64_t dh = (q % 2048) * kHour + rh;
    if (trunc_hours(dh) != ref_trunc_hours(dh)) bad++;
    n += 2;
  }
  // boundaries
  for (int64_t q : {int64_t(0), int64_t(1), (int64_t(1) << 15) - 1}) for (int64_t r : {int64_t(0), kMinute - 1}) {
    int64_t d = q * kMinute + r; if (floor_minutes_over(d, 1) != ref_floor_minutes_over(d, 1)) bad++;
  }
  // since(): saturating now - t against a 128-bit reference
  auto ref_since = [](int64_t now, int64_t t) -> int64_t {
    if (t == EVG_TIME_ZERO) return kI64Max;
    __int128 d = (__int128)now - (__int128)t;
    if (d > kI64Max) return kI64Max;
    if (d < kI64Min) return kI64Min;
    return int64_t(d);
  };
  const int64_t edges[] = {0, 1, -1, kI64Max, kI64Min, kI64Min + 1, kI64Max - 1, 1800000000000000000LL, -1800000000000000000LL};
  for (int64_t a : edges) for (int64_t b : edges) { if (since(a, b) != ref_since(a, b)) bad++; n++; }
  for (int it = 0; it < 2000000; it++) {
    int64_t a = int64_t(rnd()), b = int64_t(rnd());
    if (it & 1) { a >>= (rnd() % 40); b >>= (rnd() % 40); }
    if (since(a, b) != ref_since(a, b)) bad++;
    n++;
  }
  // single_task_value == acc_add + unit_value for a lone non-task-group task
  for (int it = 0; it < 3000000; it++) {
    evg_distro_cfg c;
    c.patch_factor = int64_t(rnd() % 120) - 10; c.patch_time_in_queue_factor = int64_t(rnd() % 120) - 10;
    c.commit_queue_factor = int64_t(rnd() % 120) - 10; c.mainline_time_in_queue_factor = int64_t(rnd() % 120) - 10;
    c.expected_runtime_factor = int64_t(rnd() % 120) - 10; c.generate_task_factor = int64_t(rnd() % 120) - 10;
    c.stepback_task_factor = int64_t(rnd() % 120) - 10; c.num_dependents_factor = double(int64_t(rnd() % 1200) - 100) / 10.0;
    c.target_time_ns = 0; c.group_versions = 0; c.includes_dependencies = 0; c.n_versions = 0; c._reserved = 0;
    const int64_t now = 1800000000000000000LL;
    int64_t qb;
    switch (rnd() % 6) { case 0: qb = EVG_TIME_ZERO; break; case 1: qb = 0; break; case 2: qb = now + int64_t(rnd() % 1000000000000ULL); break;
                         case 3: qb = now - int64_t(rnd() % (30ULL * 24 * 3600 * 1000000000ULL)); break; default: qb = now - int64_t(rnd() % (72ULL * 3600 * 1000000000ULL)); }
    const int32_t prio = int32_t(rnd() % 1200) - 100, nd = int32_t(rnd() % 50) - 5;
    const int64_t ex = (rnd() % 5 == 0) ? int64_t(rnd() % (1ULL << 58)) : int64_t(rnd() % (7200ULL * 1000000000ULL));
    const uint32_t fl = uint32_t(rnd() % 3) | (rnd() % 8 == 0 ? EVG_TF_GENERATE : 0) | (rnd() % 8 == 0 ? EVG_TF_STEPBACK : 0);
    UnitAcc a; acc_init(a); acc_add(a, now, prio, ex, qb, nd, -1, fl);
    if (unit_value(a, c, nullptr) != single_task_value(clamp_factors(c), now, prio, ex, qb, nd, fl)) bad++;
    n++;
  }
  printf("checked %ld, mismatches %ld\n", n, bad);
  return bad != 0;
}
