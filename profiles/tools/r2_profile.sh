set -x
P="python profiles/run_kernel_profile.py"
$P nn 16384 > gpurun_out/plain_nn.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ya_k_ -s 900 -c 120 --csv --log-file gpurun_out/r2_launches_mcts_nn.csv $P nn 16384 > gpurun_out/ncu_nn1.log 2>&1
$P nn 16384 > gpurun_out/plain_nn.log 2>&1 && ncu --set full --clock-control none --import-source on -k 'regex:ya_k_(mcts_select|mcts_expand_rows|forward)' -s 1101 -c 3 -f -o gpurun_out/r2_mcts_nn $P nn 16384 > gpurun_out/ncu_nn2.log 2>&1
$P uniform > gpurun_out/plain_uni.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ya_k_mcts_search_uniform -s 4 -c 1 -f -o gpurun_out/r2_mcts_uniform $P uniform > gpurun_out/ncu_uni.log 2>&1
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$B > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 200 --csv --log-file gpurun_out/r2_launches_bench.csv $B > gpurun_out/ncu_b1.log 2>&1
$B > gpurun_out/plain_bench.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ya_k_play_ply -s 150 -c 2 -f -o gpurun_out/r2_play_ply $B > gpurun_out/ncu_b2.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches*; tail -3 gpurun_out/ncu_nn2.log gpurun_out/ncu_uni.log gpurun_out/ncu_b2.log
