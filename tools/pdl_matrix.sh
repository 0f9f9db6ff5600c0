# programmatic dependent launch (MSU_PDL) x weight-gradient side stream (MSUNET_B200_WGRAD_STREAM), same box
for rep in 1 2; do for pdl in 0 1 2; do for wg in 1 0; do
echo -n "PDL=$pdl WGRAD_STREAM=$wg : "
MSU_PDL=$pdl MSUNET_B200_WGRAD_STREAM=$wg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c134-160
done; done; done
