// inst.cu -- explicit instantiation of the kernels of ONE flow-family instance (-DDYN_INST=k).
#include "adjoint_solver.cuh"

#ifndef DYN_INST
#error "compile with -DDYN_INST=<index from instances.def>"
#endif

namespace dynode {

template <int IDX> struct Inst;
#define X(IDX, FLOW, FLAGS, G, S) \
  template <> struct Inst<IDX> { static constexpr int flow = FLOW, flags = FLAGS, g = G, s = S; };
#include "instances.def"
#undef X

using I = Inst<DYN_INST>;
template cudaError_t launch_lane_solver<I::flow, I::flags, I::g, I::s, 0, MODE_SAVE>(const SolveArgs&, cudaStream_t);
template cudaError_t launch_lane_solver<I::flow, I::flags, I::g, I::s, tangent_chunk(I::flow), MODE_SAVE>(const SolveArgs&, cudaStream_t);
template cudaError_t launch_lane_solver<I::flow, I::flags, I::g, I::s, 0, MODE_SAVE_JUMPS>(const SolveArgs&, cudaStream_t);
template cudaError_t launch_lane_solver<I::flow, I::flags, I::g, I::s, tangent_chunk(I::flow), MODE_SAVE_JUMPS>(const SolveArgs&, cudaStream_t);
template cudaError_t launch_lane_solver<I::flow, I::flags, I::g, I::s, 0, MODE_LOGLIK_JUMPS>(const SolveArgs&, cudaStream_t);
template cudaError_t launch_lane_solver<I::flow, I::flags, I::g, I::s, tangent_chunk(I::flow), MODE_LOGLIK_JUMPS>(const SolveArgs&, cudaStream_t);
template cudaError_t launch_lane_solver<I::flow, I::flags, I::g, I::s, 0, MODE_LOGLIK>(const SolveArgs&, cudaStream_t);
template cudaError_t launch_lane_solver<I::flow, I::flags, I::g, I::s, tangent_chunk(I::flow), MODE_LOGLIK>(const SolveArgs&, cudaStream_t);

template cudaError_t launch_adjoint_solver<I::flow, I::flags, I::g, I::s, false>(const AdjointArgs&, cudaStream_t);
template cudaError_t launch_adjoint_solver<I::flow, I::flags, I::g, I::s, true>(const AdjointArgs&, cudaStream_t);

template <int FLOW, int FLAGS, int G, int S>
cudaError_t launch_loglik_single_direction(const SolveArgs& a, cudaStream_t stream) {
  if constexpr (tangent_chunk(FLOW) > 1) {
    return launch_lane_solver<FLOW, FLAGS, G, S, 1, MODE_LOGLIK>(a, stream);
  } else {
    return cudaErrorNotSupported;
  }
}
template cudaError_t launch_loglik_single_direction<I::flow, I::flags, I::g, I::s>(const SolveArgs&, cudaStream_t);

}  // namespace dynode
