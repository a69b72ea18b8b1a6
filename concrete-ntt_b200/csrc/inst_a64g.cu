// Instantiates every kernel of arithmetic class A64G (see arith.cuh).
#include "dispatch.hpp"
namespace cntt {
CNTT_DEFINE_CLASS(A64G)
}
