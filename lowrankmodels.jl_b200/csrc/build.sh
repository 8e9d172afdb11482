#!/bin/bash
# Builds the CUDA engine (sm_100a only) and the host-side synthetic-data helper, in-tree (see Makefile).
set -e
cd "$(dirname "$0")"
make -s -j"$(nproc)" "$@"
