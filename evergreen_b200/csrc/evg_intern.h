// evg_intern.h -- evg_intern_columns: the string work of marshalling a tick, on the host, in C++.
//
// What the shim (or soa.marshal_tasks) does with maps of strings per distro -- task-group key -> dense group id,
// version -> dense version id, dependency task id -> index of that task in the distro's queue (planner.go:431-456 files
// units by exactly these strings) -- done over packed string columns with one open-addressing table per distro and the
// distros spread over threads.  Ids are handed out in first-appearance order, like the maps the reference builds while
// it walks the queue.  Pure host code: no context, no device.
#pragma once
#include <atomic>
#include <thread>

namespace evg_intern {

struct Str { const char* p; uint32_t n; };
inline uint64_t hash_bytes(const char* p, uint32_t n) {  // eight bytes per multiply (task ids are ~100 bytes); a final mix spreads the high half
  uint64_t h = 0x9E3779B97F4A7C15ull ^ (uint64_t(n) * 0xD6E8FEB86659FD93ull);
  uint32_t i = 0;
  for (; i + 8 <= n; i += 8) {
    uint64_t w;
    memcpy(&w, p + i, 8);
    h = (h ^ w) * 0xFF51AFD7ED558CCDull;
    h = (h << 29) | (h >> 35);
  }
  uint64_t w = 0;
  if (i < n) memcpy(&w, p + i, n - i);
  h = (h ^ w) * 0xC4CEB9FE1A85EC53ull;
  h ^= h >> 32;
  h *= 0xD6E8FEB86659FD93ull;
  return h ^ (h >> 29);
}
inline Str str_at(const char* bytes, const int64_t* off, int64_t i) { return Str{bytes + off[i], uint32_t(off[i + 1] - off[i])}; }
inline bool same(const Str& a, const Str& b) { return a.n == b.n && (a.n == 0 || memcmp(a.p, b.p, a.n) == 0); }

// string -> small integer, keys owned by the caller's column; one distro at a time, reused across distros
struct Table {
  std::vector<uint64_t> hashes;
  std::vector<int64_t> rows;   // row of the key's first occurrence (the key's bytes live there)
  std::vector<int32_t> values;
  std::vector<uint32_t> gens;  // slot k is occupied iff gens[k] == gen: starting a new distro is gen++, not a memset
  uint32_t gen = 0;
  uint64_t mask = 0;
  void reset(int64_t n_keys) {
    uint64_t cap = 16;
    while (cap < uint64_t(n_keys) * 2 + 2) cap <<= 1;
    if (cap > hashes.size() || gen == 0xFFFFFFFFu) {
      hashes.assign(cap, 0); rows.assign(cap, -1); values.assign(cap, 0); gens.assign(cap, 0);
      gen = 0;
    }
    gen++;
    mask = hashes.size() - 1;
  }
  // value of `s`, inserting `fresh` when absent (*inserted tells which)
  int32_t get_or_put(const Str& s, const char* bytes, const int64_t* off, int64_t row, int32_t fresh, bool* inserted) {
    const uint64_t h = hash_bytes(s.p, s.n);
    for (uint64_t k = h & mask;; k = (k + 1) & mask) {
      if (gens[k] != gen) { gens[k] = gen; hashes[k] = h; rows[k] = row; values[k] = fresh; *inserted = true; return fresh; }
      if (hashes[k] == h && same(s, str_at(bytes, off, rows[k]))) { *inserted = false; return values[k]; }
    }
  }
  int32_t find(const Str& s, const char* bytes, const int64_t* off) const {
    const uint64_t h = hash_bytes(s.p, s.n);
    for (uint64_t k = h & mask;; k = (k + 1) & mask) {
      if (gens[k] != gen) return -1;
      if (hashes[k] == h && same(s, str_at(bytes, off, rows[k]))) return values[k];
    }
  }
};

}  // namespace evg_intern
