// psy_internal.h — shared by the translation units of libpsydiff_b200.so (not part of the public ABI).
#pragma once
#include <string>
#include <vector>

#include "../../include/psydiff_b200.h"

// sets the calling thread's psy_last_error() text and returns `code`
int psy_fail(int code, const char* fmt, ...);

// host-only bookkeeping of one finite-difference sweep (the eps[] fallback of R/modelTemplate.R:885-898, 905-919)
struct psy_sweep {
  int P = 0;
  double base_total = 0.0, base_nonfinite = 0.0;
  std::vector<double> Dval;             // [P*2]  Σ_p (f_side − f0)
  std::vector<double> eps_used;         // [P]
  std::vector<unsigned char> resolved;  // [P*2]
  std::vector<unsigned char> wanted;    // [P*2]
};

