mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,lts__t_sectors.sum,lts__t_sector_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,dram__bytes_read.sum,l1tex__t_sector_hit_rate.pct --clock-control none -k regex:sam_step_kernel --csv --log-file gpurun_out/lookup_only.csv python tools/lookup_only.py 2>&1 | tail -2
