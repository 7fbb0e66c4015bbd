#!/bin/bash
# e2e through the host-buffer ABI under different pipeline settings (bench.py --no-cpu); prints value / copy ceiling
run() { echo -n "$1 : "; env $1 timeout 300 python bench.py --steps 3 --no-cpu $2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['e2e']['value'], 'of', d['e2e']['copy_ceiling'])"; }
run "PLZ4CU_LANES=4" ""
run "PLZ4CU_LANES=6" ""
run "PLZ4CU_LANES=8" ""
run "PLZ4CU_LANES=4 PLZ4CU_CHUNK_MIB=32" ""
run "PLZ4CU_LANES=8 PLZ4CU_CHUNK_MIB=32" ""
run "PLZ4CU_LANES=4" "--e2e-parts 4"
run "PLZ4CU_LANES=4" "--e2e-parts 16"
run "PLZ4CU_LANES=4" "--e2e-parts 1"
