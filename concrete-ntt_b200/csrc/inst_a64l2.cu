// Instantiates every kernel of arithmetic class A64L2 (see arith.cuh).
#include "dispatch.hpp"
namespace cntt {
CNTT_DEFINE_CLASS(A64L2)
}
