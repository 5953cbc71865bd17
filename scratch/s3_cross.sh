#!/bin/bash
# stage-1 kernels against each other over N: scratch/s3_cross.sh  (prints one line per N and kind)
export PYTHONPATH=$PWD
for n in 128 200 256 300 400 450 500 600 696; do
  s=$(( 600000000 / (n * n) )); [ $s -gt 4096 ] && s=4096
  for k in 1 3; do
    echo -n "N=$n kind=$k S=$s: "; MCOSB_STAGE1=$k timeout 300 python scratch/sbr_time.py $n $s 2 2>&1 | tail -1 | sed 's/.*:: //'
  done
done
for n in 768 1000; do
  s=$(( 600000000 / (n * n) ))
  for k in 2 3; do
    echo -n "N=$n kind=$k S=$s: "; MCOSB_STAGE1=$k timeout 300 python scratch/sbr_time.py $n $s 2 2>&1 | tail -1 | sed 's/.*:: //'
  done
done
