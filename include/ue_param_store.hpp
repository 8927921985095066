// include/ue_param_store.hpp — name-keyed storage behind ue_*_set_* (C++ helper
// shared by the product library and the test oracle; boundary plumbing only,
// no physics).  The field list comes from ue_params.h.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "ue_params.h"

struct UeParams {
#define X(n) int64_t n;
  UE_INT_SCALARS(X)
#undef X
#define X(n) double n;
  UE_REAL_SCALARS(X)
#undef X
#define X(n) const double* n;
  UE_REAL_PLANES(X)
  UE_REAL_LINES(X)
#undef X
#define X(n) const int64_t* n;
  UE_INT_PLANES(X)
  UE_INT_LINES(X)
#undef X
};

struct UeStore {
  UeParams p;
  std::map<std::string, int64_t*> iscal;
  std::map<std::string, double*> rscal;
  std::map<std::string, const double**> rarr;
  std::map<std::string, const int64_t**> iarr;
  std::map<std::string, bool> is_plane;
  std::map<std::string, std::vector<double>> rdata;
  std::map<std::string, std::vector<int64_t>> idata;
  std::map<std::string, bool> seen;
  std::map<std::string, double> zero_only;  // UE_ZERO_REALS / UE_ZERO_INTS: checked, never used

  UeStore() {
    std::memset(&p, 0, sizeof(p));
#define X(n) iscal[#n] = &p.n; seen[#n] = false;
    UE_INT_SCALARS(X)
#undef X
#define X(n) rscal[#n] = &p.n; seen[#n] = false;
    UE_REAL_SCALARS(X)
#undef X
#define X(n) rarr[#n] = &p.n; seen[#n] = false; is_plane[#n] = true;
    UE_REAL_PLANES(X)
#undef X
#define X(n) rarr[#n] = &p.n; seen[#n] = false; is_plane[#n] = false;
    UE_REAL_LINES(X)
#undef X
#define X(n) iarr[#n] = &p.n; seen[#n] = false; is_plane[#n] = true;
    UE_INT_PLANES(X)
#undef X
#define X(n) iarr[#n] = &p.n; seen[#n] = false; is_plane[#n] = false;
    UE_INT_LINES(X)
#undef X
#define X(n) zero_only[#n] = 0.; seen[#n] = false;
    UE_ZERO_REALS(X)
    UE_ZERO_INTS(X)
#undef X
  }
  int set_int(const char* name, int64_t v) {
    auto z = zero_only.find(name);
    if (z != zero_only.end()) { z->second = (double)v; seen[name] = true; return 0; }
    auto it = iscal.find(name);
    if (it == iscal.end()) return -1;
    *it->second = v; seen[name] = true; return 0;
  }
  int set_real(const char* name, double v) {
    auto z = zero_only.find(name);
    if (z != zero_only.end()) { z->second = v; seen[name] = true; return 0; }
    auto it = rscal.find(name);
    if (it == rscal.end()) return -1;
    *it->second = v; seen[name] = true; return 0;
  }
  int set_real_array(const char* name, const double* d, int64_t n) {
    if (!name || n < 0 || (!d && n > 0)) return -1;  // (a NULL array with n > 0 is refused, not dereferenced)
    auto it = rarr.find(name);
    if (it == rarr.end()) return -1;
    auto& v = rdata[name];
    v.assign(d, d + n);
    *it->second = v.data(); seen[name] = true; return 0;
  }
  int set_int_array(const char* name, const int64_t* d, int64_t n) {
    if (!name || n < 0 || (!d && n > 0)) return -1;
    auto it = iarr.find(name);
    if (it == iarr.end()) return -1;
    auto& v = idata[name];
    v.assign(d, d + n);
    *it->second = v.data(); seen[name] = true; return 0;
  }
  // names never set; also checks plane sizes once nx, ny are known
  std::string missing() const {
    std::string s;
    for (auto& kv : seen) if (!kv.second) { s += kv.first; s += ' '; }
    return s;
  }
  // first input of the must-be-zero lists that is not zero (empty: none)
  std::string nonzero_frozen() const {
    for (auto& kv : zero_only) if (kv.second != 0.) return kv.first;
    return std::string();
  }
  // expected length of a 1-D LINE by name: x-lines nx+2, y-lines ny+2, unknown-indexed neq / 2*neq, rate tables mpe*mpd
  // (only read when istabon = 10).  -1: no fixed length.
  int64_t line_length(const std::string& n) const {
    static const char* ylines[] = {"fgtdy", "flalfgya", "yyf", "recylb", "recyrb", "alblb", "albrb", "fngxslb", "fngxsrb", "fngxlb_use", "fngxrb_use"};
    static const char* tables[] = {"wsveh", "wsveh0", "welms1", "welms2"};
    for (const char* y : ylines) if (n == y) return p.ny + 2;
    for (const char* t : tables) if (n == t) return p.istabon == 10 ? (int64_t)p.mpe * p.mpd : -1;
    if (n == "iseqalg") return p.neq;
    if (n == "igyl") return 2 * p.neq;
    return p.nx + 2;
  }
  // arrays whose length does not fit the mesh: planes must hold (nx+2)(ny+2) values, lines their length class
  std::string bad_sizes() const {
    std::string s;
    const int64_t ncell = (p.nx + 2) * (p.ny + 2);
    auto check = [&](const std::string& name, int64_t have) {
      const int64_t want = is_plane.at(name) ? ncell : line_length(name);
      if (want >= 0 && have != want) { s += name; s += '['; s += std::to_string(have); s += " != "; s += std::to_string(want); s += "] "; }
    };
    for (auto& kv : rdata) check(kv.first, (int64_t)kv.second.size());
    for (auto& kv : idata) check(kv.first, (int64_t)kv.second.size());
    return s;
  }
  int64_t len(const char* name) const {
    auto a = rdata.find(name); if (a != rdata.end()) return (int64_t)a->second.size();
    auto b = idata.find(name); if (b != idata.end()) return (int64_t)b->second.size();
    return -1;
  }
};
