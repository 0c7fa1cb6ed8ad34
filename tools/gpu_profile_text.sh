#!/bin/bash
# ncu --set full capture of one kernel of a bench workload, reduced ON THE BOX to text (summary + dynamic SASS mix):
# the .ncu-rep files are 30-50 MB each and gpurun brings back at most 64 MB.
# usage: tools/gpu_profile_text.sh TAG KERNEL_REGEX SKIP RAYS WORKLOAD
TAG=$1; KREG=$2; SKIP=$3; RAYS=$4; WL=$5
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$KREG" -s $SKIP -c 1 -f -o /tmp/prof_$TAG \
    python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_$TAG.log 2>&1
python tools/summarize_ncu.py /tmp/prof_$TAG.ncu-rep gpurun_out/prof_$TAG.txt $RAYS $WL > /dev/null 2>&1
cp profiles/ncu_figures.json gpurun_out/ncu_figures.json   # figures of this capture (source = the text file above; rename when copying into profiles/)
python tools/sass_mix.py /tmp/prof_$TAG.ncu-rep $RAYS --dump > gpurun_out/sass_$TAG.txt 2>&1
head -12 gpurun_out/prof_$TAG.txt | tail -9
