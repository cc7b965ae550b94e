// Shim over the reference's own alias-table builder, compiled from /root/reference/src/util/AliasTable.h where it lies
// (oracle/ref/Makefile).  TEST INFRASTRUCTURE: pins rh_build_alias_table (host/Scene.cpp buildAliasTable) against the reference.
#include <cstdint>
#include <cstring>
#include <vector>
#include "AliasTable.h"   // the reference's header (-I$(REF)/src/util)

extern "C" __attribute__((visibility("default")))
void ref_build_alias_table(const float* power, uint32_t n, void* out /* (n + 1) x {float prob, uint32 failId} */) {
	DiscreteSampler1D<float> sampler(std::vector<float>(power, power + n));   // reference src/Scene.cpp:316 builds it the same way
	static_assert(sizeof(BinomialDistrib<float>) == 8, "layout of layouts.glsl:85-88");
	std::memcpy(out, sampler.binomDistribs.data(), sampler.binomDistribs.size() * sizeof(BinomialDistrib<float>));
}
