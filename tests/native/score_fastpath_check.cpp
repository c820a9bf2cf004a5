// Host check: the integer fast paths of evg_score.cuh equal the literal FP64 formulas.
#include <cstdio>
#include <initializer_list>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include "evergreen_b200/csrc/evg_score.cuh"
using namespace evg;
static int64_t ref_floor_minutes_over(int64_t d, int64_t n) { return int64_t(std::floor((double(d / kMinute) + double(d % kMinute) / (60.0 * 1e9)) / double(n))); }
static int64_t ref_trunc_hours(int64_t d) { return int64_t(double(d / kHour) + double(d % kHour) / (3600.0 * 1e9)); }
int main() {
  uint64_t x = 88172645463325252ULL; long bad = 0; long n = 0; long fast_n = 0; long n32 = 0; long n32t = 0; long nbad0 = 0;
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
    c.stepback_task_factor = int64_t(rnd() % 120) - 10; c.num_dependents_factor = (rnd() % 4) ? double(int64_t(rnd() % 120) - 10) : double(int64_t(rnd() % 1200) - 100) / 10.0;
    c.target_time_ns = 0; c.group_versions = 0; c.includes_dependencies = 0; c.n_versions = 0; c._reserved = 0;
    const int64_t now = 1800000000000000000LL;
    int64_t qb;
    switch (rnd() % 6) { case 0: qb = EVG_TIME_ZERO; break; case 1: qb = 0; break; case 2: qb = now + int64_t(rnd() % 1000000000000ULL); break;
                         case 3: qb = now - int64_t(rnd() % (30ULL * 24 * 3600 * 1000000000ULL)); break; default: qb = now - int64_t(rnd() % (72ULL * 3600 * 1000000000ULL)); }
    const int32_t prio = int32_t(rnd() % 1200) - 100, nd = int32_t(rnd() % 50) - 5;
    const int64_t ex = (rnd() % 5 == 0) ? int64_t(rnd() % (1ULL << 58)) : int64_t(rnd() % (7200ULL * 1000000000ULL));
    const uint32_t fl = uint32_t(rnd() % 3) | (rnd() % 8 == 0 ? EVG_TF_GENERATE : 0) | (rnd() % 8 == 0 ? EVG_TF_STEPBACK : 0);
    UnitAcc a; acc_init(a); acc_add(a, now, prio, ex, qb, nd, -1, fl);
    const PlannerFactors pf = clamp_factors(c);
    const int64_t want = unit_value(a, c, nullptr);
    if (want != single_task_value(pf, now, prio, ex, qb, nd, fl)) bad++;
    // the straight-line form, wherever its domain test admits the inputs
    if (pf.nd_int != 0 && score_fast_domain(now, ex, qb)) {
      fast_n++;
      if (want != single_task_value_fast(pf, now, prio, ex, qb, nd, fl)) bad++;
    }
    const Factors32 f32 = factors32(pf, now);
    if (f32.ok && score32_domain(now, prio, nd, ex, qb)) {
      n32++;
      if (uint64_t(want) != single_task_value32(f32, now, prio, ex, qb, nd, fl)) bad++;
    }
    // the tabulated-NumDependents form admits fractional factors
    if (f32.ok_base && score32_domain_nd(now, prio, ex, qb) && nd < kNdTable) {
      const int64_t e = nd_table_entry(pf, nd > 0 ? nd : 0);
      if (e >= 0 && e < int64_t(kNdTermLimit)) {
        n32t++;
        if (uint64_t(want) != single_task_value32_nd(f32, now, prio, ex, qb, uint32_t(e), fl)) bad++;
      }
    }
    // the kernels' OR-form of the domain, also with a clock near the epoch and bases before it
    {
      const int64_t now2 = (it & 1) ? now : int64_t(rnd() % (1ULL << 52));
      int64_t qb2 = qb;
      if (it % 7 == 0) qb2 = now2 - int64_t(rnd() % (1ULL << 51)) + (int64_t(1) << 49);  // straddles 0 and the 2^50 limit
      if (it % 11 == 0) qb2 = int64_t(rnd());
      const Factors32 g32 = factors32(pf, now2);
      if (g32.ok_base && nd < kNdTable) {
        const int64_t e = nd_table_entry(pf, nd > 0 ? nd : 0);
        const uint32_t term = (e >= 0 && e < int64_t(kNdTermLimit)) ? uint32_t(e) : 0xFFFFFFFFu;
        if (score32_bad(now2, prio, ex, qb2, term) == 0u) {
          nbad0++;
          UnitAcc a2; acc_init(a2); acc_add(a2, now2, prio, ex, qb2, nd, -1, fl);
          if (uint64_t(unit_value(a2, c, nullptr)) != single_task_value32_nd(g32, now2, prio, ex, qb2, term, fl)) bad++;
        }
      }
    }
    n++;
  }
  // the straight-line form at the edges of its domain (week boundary, limit - 1, zero basis, huge factors)
  {
    const int64_t now = 1800000000000000000LL;
    const int64_t lim = int64_t(kFastLimit);
    const int64_t tiqs[] = {0, 1, kMinute - 1, kMinute, kHour - 1, kHour, kWeek - kHour, kWeek - 1, kWeek, kWeek + 1, lim - 1};
    const int64_t exps[] = {0, 1, kMinute - 1, kMinute, 90 * kMinute, lim - 1};
    const int64_t facs[] = {0, 1, 7, 100, 2147483647LL, 9007199254740993LL, -5};
    const double ndf[] = {0.0, 1.0, 2.0, 1048575.0, 1048576.0, 2.5, -1.0};
    for (int64_t tq : tiqs) for (int64_t ex : exps) for (int64_t fc : facs) for (double nf : ndf) for (uint32_t fl = 0; fl < 3; fl++)
      for (int extra = 0; extra < 4; extra++) for (int zero = 0; zero < 2; zero++) {
        evg_distro_cfg c;
        c.patch_factor = fc; c.patch_time_in_queue_factor = fc + 1; c.commit_queue_factor = fc; c.mainline_time_in_queue_factor = fc + 2;
        c.expected_runtime_factor = fc; c.generate_task_factor = fc; c.stepback_task_factor = fc + 3; c.num_dependents_factor = nf;
        c.target_time_ns = 0; c.group_versions = 0; c.includes_dependencies = 0; c.n_versions = 0; c._reserved = 0;
        const uint32_t f2 = fl | ((extra & 1) ? EVG_TF_GENERATE : 0) | ((extra & 2) ? EVG_TF_STEPBACK : 0);
        const int64_t qb = zero ? EVG_TIME_ZERO : now - tq;
        const int32_t prio = int32_t(tq % 3 == 0 ? 2147483647 : 17), nd = int32_t(ex % 2 == 0 ? 2147483647 : 3);
        UnitAcc a; acc_init(a); acc_add(a, now, prio, ex, qb, nd, -1, f2);
        const PlannerFactors pf = clamp_factors(c);
        if (unit_value(a, c, nullptr) != single_task_value(pf, now, prio, ex, qb, nd, f2)) bad++;
        if (pf.nd_int != 0 && score_fast_domain(now, ex, qb)) {
          fast_n++;
          if (unit_value(a, c, nullptr) != single_task_value_fast(pf, now, prio, ex, qb, nd, f2)) bad++;
        }
        // the 32-bit form at the edges of ITS domain: factors up to 2^14 - 1, priority / dependents up to 2^15 - 1
        for (int big = 0; big < 2; big++) {
          evg_distro_cfg c2 = c;
          if (big) {
            c2.patch_factor = c2.patch_time_in_queue_factor = c2.commit_queue_factor = c2.mainline_time_in_queue_factor =
                c2.expected_runtime_factor = c2.generate_task_factor = c2.stepback_task_factor = kFactor32Limit - 1;
            c2.num_dependents_factor = double(kFactor32Limit - 1);
          }
          const int32_t p2 = big ? int32_t(kTask32Limit) - 1 : int32_t(tq % 5) - 1, d2 = big ? int32_t(kTask32Limit) - 1 : int32_t(ex % 7) - 1;
          const PlannerFactors pf2 = clamp_factors(c2);
          const Factors32 f32 = factors32(pf2, now);
          if (f32.ok && score32_domain(now, p2, d2, ex, qb)) {
            UnitAcc a2; acc_init(a2); acc_add(a2, now, p2, ex, qb, d2, -1, f2);
            n32++;
            if (uint64_t(unit_value(a2, c2, nullptr)) != single_task_value32(f32, now, p2, ex, qb, d2, f2)) bad++;
          }
        }
        n++;
      }
    // outside the domain the test must say so
    if (score_fast_domain(now, lim, 0 + now - 1) || score_fast_domain(now, 0, now - lim) || score_fast_domain(now, -1, now) ||
        score_fast_domain(now, 0, now + 1) || score_fast_domain(now, 0, -5)) bad++;
  }
  if (fast_n < 1000000) bad++;  // the straight-line form must actually have been exercised
  if (n32 < 500000 || n32t < 500000 || nbad0 < 500000) bad++;      // and so must the 32-bit forms
  {  // out of the 32-bit domain: big factor, big priority, big dependents, negative clock
    evg_distro_cfg c; memset(&c, 0, sizeof(c));
    c.patch_factor = kFactor32Limit;
    if (factors32(clamp_factors(c), 5).ok) bad++;
    c.patch_factor = 3; c.num_dependents_factor = 2.5;
    if (factors32(clamp_factors(c), 5).ok) bad++;
    c.num_dependents_factor = 0.0;
    if (!factors32(clamp_factors(c), 5).ok || factors32(clamp_factors(c), -5).ok) bad++;
    const int64_t now = 1800000000000000000LL;
    if (score32_domain(now, int32_t(kTask32Limit), 0, 0, now) || score32_domain(now, 0, int32_t(kTask32Limit), 0, now) ||
        !score32_domain(now, -7, -7, 0, now)) bad++;
  }
  printf("checked %ld (%ld through the straight-line form, %ld through the 32-bit form), mismatches %ld\n", n, fast_n, n32, bad);
  return bad != 0;
}
