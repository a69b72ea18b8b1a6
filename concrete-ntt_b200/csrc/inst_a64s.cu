// Instantiates every kernel of arithmetic class A64S (see arith.cuh).
#include "dispatch.hpp"
namespace cntt {
CNTT_DEFINE_CLASS(A64S)
}
