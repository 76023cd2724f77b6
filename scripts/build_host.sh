#!/usr/bin/env bash
# Builds the C++ host CLI (links the C-ABI library; runs only where a B200 is present,
# `--selftest` checks the pure host functions anywhere).  STB_HOST_CXXFLAGS adds compiler flags, e.g.
# "-fsanitize=address,undefined -g" to run the CPU tests of tests/test_host_cpp.py under the sanitizers.
set -euo pipefail
cd "$(dirname "$0")/.."
/usr/bin/g++ -std=c++17 -O2 -Wall -Wextra ${STB_HOST_CXXFLAGS:-} -o semtools_b200/lib/semtools_b200_search \
  semtools_b200/host/semtools_search_main.cpp semtools_b200/host/semtools_host.cpp semtools_b200/host/semtools_store.cpp \
  semtools_b200/host/semtools_tokenizer.cpp \
  -Lsemtools_b200/lib -lsemtools_b200 -Wl,-rpath,'$ORIGIN'
/usr/bin/g++ -std=c++17 -O2 -Wall -Wextra ${STB_HOST_CXXFLAGS:-} -o semtools_b200/lib/semtools_b200_workspace \
  semtools_b200/host/semtools_workspace_main.cpp semtools_b200/host/semtools_store.cpp semtools_b200/host/semtools_host.cpp \
  -Lsemtools_b200/lib -lsemtools_b200 -Wl,-rpath,'$ORIGIN'
