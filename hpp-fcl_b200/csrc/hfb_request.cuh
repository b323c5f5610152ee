// Request decoding: what GJKSolver::set(request) derives from a DistanceRequest /
// CollisionRequest (include/hpp/fcl/narrowphase/narrowphase.h:162-190, 214-244).
#pragma once
#include "hfb_pair.cuh"

namespace hfb {

inline void solver_from_query(const hfb_query_request& q, SolverP& P) {
  P.gjk.tolerance = q.gjk_tolerance;
  P.gjk.max_iterations = q.gjk_max_iterations;
  P.gjk.variant = q.gjk_variant;
  P.gjk.criterion = q.gjk_convergence_criterion;
  P.gjk.criterion_type = q.gjk_convergence_criterion_type;
  P.epa.tolerance = q.epa_tolerance;
  P.epa.max_iterations = q.epa_max_iterations;
  P.initial_guess = q.gjk_initial_guess;
}
inline SolverP solver_from_distance_request(const hfb_distance_request& r) {
  SolverP P;
  solver_from_query(r.q, P);
  P.gjk.distance_upper_bound = DBL_MAX;  // narrowphase.h:175
  P.compute_penetration = r.enable_signed_distance != 0;
  return P;
}
inline SolverP solver_from_collision_request(const hfb_collision_request& r) {
  SolverP P;
  solver_from_query(r.q, P);
  const double ub = r.distance_upper_bound > r.security_margin ? r.distance_upper_bound : r.security_margin;
  P.gjk.distance_upper_bound = ub > 0. ? ub : 0.;  // narrowphase.h:228-229
  P.compute_penetration = (r.enable_contact != 0) || (r.security_margin < 0);  // shape_shape_func.h:141-142
  return P;
}
// argument validation shared by every entry point (the reference throws
// std::invalid_argument; here an error code)
inline int validate_query(const hfb_query_request& q) {
  if (!(q.gjk_tolerance > 0) || !(q.epa_tolerance > 0)) return HFB_ERR_INVALID_ARGUMENT;  // gjk.cpp:62
  if (q.epa_max_iterations > HFB_EPA_CAP_IT) return HFB_ERR_INVALID_ARGUMENT;
  if (q.gjk_variant < 0 || q.gjk_variant > 2) return HFB_ERR_INVALID_ARGUMENT;
  if (q.gjk_convergence_criterion < 0 || q.gjk_convergence_criterion > 2) return HFB_ERR_INVALID_ARGUMENT;
  if (q.gjk_initial_guess < 0 || q.gjk_initial_guess > 2) return HFB_ERR_INVALID_ARGUMENT;
  if (q.gjk_convergence_criterion_type < 0 || q.gjk_convergence_criterion_type > 1) return HFB_ERR_INVALID_ARGUMENT;
  return HFB_OK;
}

}  // namespace hfb
