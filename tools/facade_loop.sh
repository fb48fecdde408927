#!/bin/bash
# build the C++ facade test once and run it N times (flakiness check of the threaded / peer-exchange paths)
set -e
N=${1:-20}
LIB=qrkit_b200/lib/libqrkit_b200.so
g++ -std=c++17 -O1 -I include tests/cpp/test_facade.cpp -o /tmp/test_facade $LIB -Wl,-rpath,$PWD/qrkit_b200/lib
fail=0
for i in $(seq $N); do
  if ! /tmp/test_facade > /tmp/facade_out.txt 2>&1; then fail=$((fail+1)); echo "run $i:"; grep -v "^  ok\|^ok" /tmp/facade_out.txt | tail -5; fi
done
echo "facade: $fail of $N runs failed"
