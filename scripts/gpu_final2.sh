#!/bin/bash
# end-of-round check: smoke(), the whole GPU suite, both bench arms, the launch list and one --set full capture of k_fdr
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/r4_final_smoke.log
bash scripts/gpu_round.sh r4_final
