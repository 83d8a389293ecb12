for rep in 1 2 3; do for v in A B; do
 echo "lib=$v rep=$rep" >> gpurun_out/sweep.log
 DNAB_LIB=$PWD/ab/lib$v.so timeout 120 python tools_probe.py cfg2 33 >> gpurun_out/sweep.log 2>&1
done; done
