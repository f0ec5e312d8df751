#!/bin/bash
# Runs the CPU tier's native-host tests (tests/test_wire_native_cpu.py, tests/test_native_host_cpu.py) with libbzk's host sources
# built under AddressSanitizer: csrc/mpn_host.cu, mpn_wire.cu, mpn_prover.cu, mpn_circuit.cu, poseidon_host.cu compiled with g++
# over the fake CUDA runtime of tests/hostshim (see tests/conftest.py::hostmpn).  Last run: 16 passed, no report.
set -e
cd "$(dirname "$0")/.."
S=tests/hostshim/_mpn_shim.so
[ -f $S ] && cp $S /tmp/_mpn_shim_plain.so
g++ -O1 -g -fsanitize=address -fno-omit-frame-pointer -std=c++17 -shared -fPIC -w -x c++ -include tests/hostshim/fake_cuda_pre.h \
    -I bazuka_b200/csrc -I /usr/local/cuda/include bazuka_b200/csrc/{mpn_host,mpn_wire,mpn_prover,mpn_circuit,poseidon_host}.cu \
    tests/hostshim/mpn_shim.cpp -x none -Wl,-Bsymbolic bazuka_b200/libbzk.so -Wl,-rpath,$PWD/bazuka_b200 -o $S
touch $S
ASAN_OPTIONS=detect_leaks=0 LD_PRELOAD=$(gcc -print-file-name=libasan.so) python -m pytest tests/test_wire_native_cpu.py tests/test_native_host_cpu.py -x -q
rm -f $S
[ -f /tmp/_mpn_shim_plain.so ] && cp /tmp/_mpn_shim_plain.so $S && touch $S
