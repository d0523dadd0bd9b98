# round 2, second half: captures of the CTA-pair forward kernels (one tile per CTA at 16,384 leaves, two tiles at 37,888)
set -x
P="python profiles/run_kernel_profile.py"
$P nn 16384 > gpurun_out/plain_nn.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ya_k_ -s 900 -c 120 --csv --log-file gpurun_out/r2b_launches_mcts_nn.csv $P nn 16384 > gpurun_out/ncu_nn1.log 2>&1
$P nn 16384 > gpurun_out/plain_nn.log 2>&1 && ncu --set full --clock-control none --import-source on -k 'regex:ya_k_(mcts_select|mcts_expand_rows|forward)' -s 1101 -c 3 -f -o gpurun_out/r2b_mcts_nn $P nn 16384 > gpurun_out/ncu_nn2.log 2>&1
$P nn 37888 > gpurun_out/plain_nn2.log 2>&1 && ncu --set full --clock-control none --import-source on -k 'regex:ya_k_forward2' -s 400 -c 1 -f -o gpurun_out/r2b_forward2 $P nn 37888 > gpurun_out/ncu_nn3.log 2>&1
python profiles/tools/forward_timeline.py 16384 --tiles=1 > gpurun_out/r2b_tl_1tile_16384.txt 2>&1
python profiles/tools/forward_timeline.py 16384 --tiles=1 --scatter=bid > gpurun_out/r2b_tl_1tile_16384_bid.txt 2>&1
python profiles/tools/forward_timeline.py 16384 --tiles=1 --scatter=score > gpurun_out/r2b_tl_1tile_16384_score.txt 2>&1
python profiles/tools/forward_timeline.py 37888 --tiles=2 > gpurun_out/r2b_tl_2tile_37888.txt 2>&1
python profiles/tools/forward_timeline.py 37888 --tiles=2 --scatter=bid > gpurun_out/r2b_tl_2tile_37888_bid.txt 2>&1
python profiles/tools/forward_timeline.py 37888 --tiles=1 > gpurun_out/r2b_tl_1tile_37888.txt 2>&1
profiles/tools/_dbg/pipe_rates > gpurun_out/r2b_pipe_rates.txt 2>&1
profiles/tools/_dbg/store_pattern > gpurun_out/r2b_store_pattern.txt 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu_nn2.log gpurun_out/ncu_nn3.log
