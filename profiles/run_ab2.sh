out=gpurun_out/r2b; mkdir -p $out
for lib in $(ls profiles/ab/*.so); do
  DDGI_LIB=$lib timeout 300 python profiles/ab_kernel.py field_32,cave_128 2 16 >> $out/ab.txt 2>&1
done
cat $out/ab.txt
