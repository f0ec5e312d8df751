// TEST INFRASTRUCTURE (CPU tier only).  Force-included when the HOST parts of libbzk (csrc/mpn_host.cu, csrc/mpn_wire.cu,
// csrc/poseidon_host.cu) are compiled with g++ for tests/hostshim/_mpn_shim.so: the one device intrinsic that common.cuh's
// templates name outside of any instantiation.
#pragma once
template <class T>
static inline T __ldg(const T *p) { return *p; }
