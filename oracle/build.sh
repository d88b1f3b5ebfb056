#!/bin/bash
# Builds the CPU oracle (test infrastructure) into oracle/liboracle.so.
# zstd: linked against the image's libzstd.so.1 (prototypes declared in vlo_util.h; no zstd.h in the image).
set -e
cd "$(dirname "$0")"
g++ -std=c++17 -O2 -fPIC -shared -Wall -Wno-unused-function -pthread vlo_api.cpp -o liboracle.so -l:libzstd.so.1
echo built oracle/liboracle.so
