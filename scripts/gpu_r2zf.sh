#!/bin/bash
# round 2, session 5: e2e step time distribution (66 steps each) at chunk 500 / 768, gallery-first copy order, and 768 with the old order
mkdir -p gpurun_out
MADE_DIAG_STEPS=66 MADE_DIAG_CHUNKS=500,768,500,768 timeout 200 python scripts/diag_e2e.py dma 2>&1 | grep "after 6" | tee gpurun_out/e2e_dist.log
MADE_PRIME_GALLERY=0 MADE_DIAG_STEPS=66 MADE_DIAG_CHUNKS=768 timeout 100 python scripts/diag_e2e.py dma 2>&1 | grep "after 6" | sed 's/^/prime=0 /' | tee -a gpurun_out/e2e_dist.log
