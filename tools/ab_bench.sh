# gpurun -- 'bash tools/ab_bench.sh': device-resident throughput of several builds of the library on ONE box, alternating
# (box-to-box differences are +-1.5 %, larger than most single changes).  Variants: any other builds of the library dropped into groundsegmentationrosnode_b200/variants/ (git-ignored; selected with GS_B200_LIB)
D=gpurun_out/ab; mkdir -p $D; : > $D/ab.txt
for round in 1 2; do
  for lib in groundsegmentationrosnode_b200/libgs_b200.so groundsegmentationrosnode_b200/variants/*.so; do
    v=$(GS_B200_LIB=$PWD/$lib python bench.py --steps 5 --warmup 3 --frames ${AB_FRAMES:-2048} --no-cpu-baseline --no-parity --e2e-frames 64 2>/dev/null | python -c "import sys,json; print(json.loads(sys.stdin.read().strip().splitlines()[-1])['value'])")
    echo "$round $lib $v" >> $D/ab.txt
  done
done
cat $D/ab.txt
