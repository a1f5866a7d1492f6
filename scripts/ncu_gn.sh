T=/tmp/vfprof; mkdir -p $T
VF_PDL=0 timeout 600 ncu --set full --import-source on -k regex:"gn_bwd" -s 6 -c 4 -o $T/gnb -f python scripts/train_probe.py 28 noprof > gpurun_out/ncu_gn.log 2>&1
python scripts/ncu_summary.py $T/gnb.ncu-rep
for i in 0 1; do ncu -i $T/gnb.ncu-rep --page details --kernel-id ::regex:gn_bwd:$((i+1)) 2>/dev/null | grep -E "gn_bwd|Warp Cycles Per Issued|Stall|Issue Slots Busy|Executed Ipc|No Eligible|Eligible Warps|L1/TEX Hit|L2 Hit|Sectors/Req|Local|Bank conflict|DRAM Throughput|Max Bandwidth|Theoretical Occ|Achieved Occ|Est. Speedup|uncoalesced|Avg. Active Threads" | head -40; done
python scripts/ncu_top.py $T/gnb.ncu-rep 12 | tail -16
