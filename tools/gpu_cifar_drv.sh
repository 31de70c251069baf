#!/bin/bash
cd tests/cpp
for sd in 3 4 5; do
echo "== seed $sd batch"; ./cifar_train --synthetic 6000 --batch 10 --epochs 6 --lr 0.005 --examples 500 --seed $sd | grep -E "correct" | tail -1
echo "== seed $sd sgd"; ./cifar_train --synthetic 5000 --epochs 3 --lr 0.02 --examples 200 --seed $sd | grep -E "correct" | tail -1
done
