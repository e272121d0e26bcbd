#!/bin/bash
# Runs GPU-tier test files against the CPU emulation of the complete library (no GPU needed; slow: one OS thread plays one CUDA
# thread).  The ctypes wrapper loads whatever MVO_LIB names, so the tests themselves are unchanged.
#   tools/replay_gpu_tests_on_emulation.sh tests/test_homography_gpu.py [more files / pytest options]
# Useful before spending GPU minutes on a kernel change.  Files that hand device pointers to the library (torch CUDA tensors) or run
# the C++ demo binaries (linked against the real libmvo.so) cannot be replayed this way.
set -e
cd "$(dirname "$0")/.."
EMU_DIR=${MVO_EMU_DIR:-/tmp/mvo_emu_full}
python -c "import sys; sys.path.insert(0, 'tests'); import emu_build; print(emu_build.build('$EMU_DIR', emu_build.ALL_UNITS))"
MVO_LIB=$EMU_DIR/libmvo_emu.so MVO_TEST_TIMEOUT_SCALE=${MVO_TEST_TIMEOUT_SCALE:-30} python -m pytest "$@" -m gpu --runxfail -q -p no:cacheprovider
