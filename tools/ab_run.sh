#!/bin/bash
# A/B harness: runs a command once per library variant in ab/ (each copied over libb2mj.so), two rounds
for round in 1 2; do
  for f in ab/lib_*.so; do
    cp "$f" mujoco-template_b200/libb2mj.so
    echo -n "$(basename $f) r$round: "
    "$@" 2>&1 | tail -1
  done
done
