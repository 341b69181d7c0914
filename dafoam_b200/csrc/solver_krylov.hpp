#pragma once
namespace dab {
inline void Solver::calcPC() { throw Error("not implemented"); }
inline int Solver::solveLinearEqn(const double*, double*, KspStats&) { throw Error("not implemented"); }
}
