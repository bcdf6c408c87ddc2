#!/bin/bash
# Local wrapper: rebuild the in-tree library (so the snapshot carries a current .so), then run a command on the B200 box.
#   scripts/gpu.sh [--gpus N] [--timeout S] -- '<command>'
cd "$(dirname "$0")/.." || exit 1
python -c "import __graft_entry__ as g; g.build()" || exit 1
exec /usr/local/graft/bin/gpurun "$@"
