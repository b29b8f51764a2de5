#!/bin/bash
# A/B of the working tree against a checkout of HEAD in _head/ (its library built from HEAD's sources): interleaved runs of
# tools/ab.py's child, wavefront and megakernel, same box.   tools/ab_head.sh [reps] [spp]
# Prepare (here, before gpurun):  rm -rf _head && mkdir _head && git archive HEAD | tar -x -C _head && (cd _head && python -m path_tracer_python_b200.build --only=strict)
# (_head/ is git-ignored but travels with the gpurun snapshot.)
REPS=${1:-3}; SPP=${2:-256}
for i in $(seq $REPS); do
  echo -n "new:  "; AB_CHILD=1 SPP=$SPP python tools/ab.py
  echo -n "head: "; (cd _head && AB_CHILD=1 SPP=$SPP python tools/ab.py)
done
