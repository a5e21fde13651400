out=gpurun_out/r2g; mkdir -p $out
for lib in "" $(ls profiles/ab/*.so); do DDGI_LIB=$lib timeout 300 python profiles/ab_pixel.py field_32,cave_128,cave_64 >> $out/abpx.txt 2>&1; done
timeout 600 python -m pytest tests -m gpu -q -x -k "frame or cave or golden or modes or edge or fuzz or baseline or octahedral" > $out/pytest.log 2>&1; tail -3 $out/pytest.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:render_frame_kernel -s 2 -c 1 -f -o $out/prof_px python profiles/ab_pixel.py field_32 3 > $out/ncu_px.log 2>&1
cat $out/abpx.txt
