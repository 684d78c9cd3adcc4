#!/bin/sh
# Build the CPU unit-test harness (test infrastructure; see hostsim.cpp header).
set -e
cd "$(dirname "$0")"
g++ -O2 -g -std=c++17 -fPIC -shared -ffp-contract=off -o libfid_hostsim.so hostsim.cpp
