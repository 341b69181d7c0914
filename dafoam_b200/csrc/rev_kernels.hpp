// STUB (replaced below once the forward is verified)
#pragma once
#include "views.hpp"
namespace dab {
struct RevA { MeshView m; Params q; StateView s; RecordView r; AdjView a; const double* x; DAB_HD void operator()(int) const {} };
struct RevB { MeshView m; Params q; StateView s; RecordView r; AdjView a; const double* x; double* y; DAB_HD void operator()(int) const {} };
struct RevC { MeshView m; Params q; StateView s; RecordView r; AdjView a; const double* x; double* y; int functionMode = 0; DAB_HD void operator()(int) const {} };
struct ForceSpec { unsigned mask; double dir[3]; double scale; };
struct ForceFwd { MeshView m; Params q; StateView s; RecordView r; ForceSpec fs; double* out; DAB_HD void operator()(int) const {} };
struct ForceRevA { MeshView m; Params q; StateView s; RecordView r; AdjView a; ForceSpec fs; double seed; DAB_HD void operator()(int) const {} };
}
