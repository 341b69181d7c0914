#pragma once
#include "backend.hpp"
namespace dab {
struct KspStats { int iterations = 0, reason = 0; double r0 = 0, rn = 0, solveSec = 0, pcSec = 0; int nMatvec = 0; };
struct Krylov { bool pcValid = false; };
}
