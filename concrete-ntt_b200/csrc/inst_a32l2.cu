// Instantiates every kernel of arithmetic class A32L2 (see arith.cuh).
#include "dispatch.hpp"
namespace cntt {
CNTT_DEFINE_CLASS(A32L2)
}
