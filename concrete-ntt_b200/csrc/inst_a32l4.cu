// Instantiates every kernel of arithmetic class A32L4 (see arith.cuh).
#include "dispatch.hpp"
namespace cntt {
CNTT_DEFINE_CLASS(A32L4)
}
