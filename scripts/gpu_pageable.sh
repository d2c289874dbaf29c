#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; nproc
for t in 2 4 8 12 16; do B200SOXR_COPY_THREADS=$t python scratch/gpu_pageable.py; done 2>&1 | tee $OUT/r2_pageable.log
for c in 4 8 32 64; do B200SOXR_CHUNK_MB=$c python scratch/gpu_pageable.py; done 2>&1 | tee -a $OUT/r2_pageable.log
B200SOXR_NO_HOST_STAGING=1 python scratch/gpu_pageable.py 2>&1 | sed 's/^/driver staging (no pinned slots): /' | tee -a $OUT/r2_pageable.log
