for mb in 40 48; do
for lib in lifo; do
export MCF_L2_PERSIST_MB=$mb
echo "persist_mb=$mb $lib"
bash tools/variant_bench.sh variants/lib_384_1_$lib.so
MCF_LIB_PATH=$PWD/variants/lib_384_1_$lib.so ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "regex:^k_grid$" --launch-skip 4 --launch-count 1 --csv --log-file gpurun_out/dram_${lib}_$mb.csv python bench.py --rows 4096 --band-cols 1024 --win-days 10 --steps 2 --warmup 3 --no-cpu --e2e-rows 256 --e2e-cols 256 --e2e-hours 48 > /dev/null 2>&1
grep -o '"dram__bytes[^"]*","byte","[0-9]*"' gpurun_out/dram_${lib}_$mb.csv | tr '\n' ' '; echo
done; done
