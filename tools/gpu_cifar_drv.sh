#!/bin/bash
cd tests/cpp
for a in "--batch 50 --epochs 10 --lr 0.0005" "--batch 50 --epochs 10 --lr 0.001" "--batch 50 --epochs 10 --lr 0.003" "--batch 10 --epochs 6 --lr 0.005"; do
  echo "== $a"; ./cifar_train --synthetic 6000 $a --examples 500 | grep -E "correct" | tail -1
done
