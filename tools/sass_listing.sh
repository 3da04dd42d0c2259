#!/bin/bash
# Regenerates profiles/sass_{gemm_tc,decoder_fused,planesweep,conv3d_tc}.txt from the objects of the last build:
# mnemonic counts of the tensor / copy / barrier / reduction instructions, then those instructions per kernel.
#   python 3dvnet_b200/build.py && bash tools/sass_listing.sh
cd "$(dirname "$0")/.."
PAT='UTCHMMA|UTCBAR|UTCATOMSWS|LDTM|STTM|UBLKCP|UTMALDG|ACQBULK|SYNCS|ELECT|REDUX|REDG|RED\.|ATOMG'
for k in gemm_tc decoder_fused planesweep conv3d_tc; do
    o=3dvnet_b200/build/$k.o
    [ -f "$o" ] || { echo "missing $o"; exit 1; }
    out=profiles/sass_$k.txt
    cuobjdump -sass "$o" > /tmp/sass_$k.txt
    {
        echo "# cuobjdump -sass $o (sm_100a), round 2 - mnemonic counts, then the tensor / copy / barrier instructions of every kernel"
        echo "#"
        grep -oE "($PAT)[A-Z0-9_.]*" /tmp/sass_$k.txt | sort | uniq -c | sort -rn | sed 's/^/# /'
        echo "#"
        grep -E "Function :|$PAT" /tmp/sass_$k.txt
    } > "$out"
    echo "$out: $(wc -l < "$out") lines"
done
