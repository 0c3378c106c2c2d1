#!/bin/bash
# round 2, session 5: e2e step against the ingest chunk size with the gallery-first copy order
mkdir -p gpurun_out
MADE_DIAG_CHUNKS=256,384,512,768,1000 timeout 300 python scripts/diag_e2e.py dma 2>&1 | grep -v "^$" | awk '{ if ($0 ~ /enqueue/) print; else { n=split($0,a," "); s=0; c=0; for(i=n-19;i<=n;i++){s+=a[i];c++}; print a[1],a[2],a[3],"mean of last 20:", s/c } }' | tee gpurun_out/e2e_chunks.log
