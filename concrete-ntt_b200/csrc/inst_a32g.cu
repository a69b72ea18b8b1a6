// Instantiates every kernel of arithmetic class A32G (see arith.cuh).
#include "dispatch.hpp"
namespace cntt {
CNTT_DEFINE_CLASS(A32G)
}
