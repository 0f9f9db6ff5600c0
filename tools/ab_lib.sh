# A/B of two builds on ONE box: the in-tree library against tools/build/lib_old.so (boxes differ by +-2 %, so never compare across calls)
L=semantic_segmentation_of_stylegan2_artifacts_b200/libmsunet_sm100.so
cp $L /tmp/new.so
for i in 1 2; do
cp /tmp/new.so $L; echo NEW; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
cp tools/build/lib_old.so $L; echo OLD; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
done
cp /tmp/new.so $L
