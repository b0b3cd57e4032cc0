#!/bin/bash
# SURVEY.md section 7 step 0(i) / VERDICT item 2: is the upstream NSGT package obtainable on the GPU box?  Records the outcome.
out=gpurun_out/cqt_pytorch_probe.txt
mkdir -p gpurun_out
{
  echo "== date: $(date -u)"; echo "== host: $(hostname)"; echo "== python: $(python -V 2>&1)"
  echo "== import cqt_pytorch"; python -c "import cqt_pytorch, sys; print('FOUND', cqt_pytorch.__file__)" 2>&1 | tail -2
  echo "== pip show cqt-pytorch"; python -m pip show cqt-pytorch 2>&1 | tail -3
  echo "== find on disk"; find / -iname "cqt_pytorch*" -not -path "/proc/*" 2>/dev/null | head
  echo "== wheelhouse"; ls /opt/wheelhouse 2>/dev/null | grep -i cqt
  echo "== pip download cqt-pytorch==0.0.4 (10 s timeout)"
  timeout 40 python -m pip download --no-deps --timeout 5 --retries 0 -d /tmp/cqt_dl cqt-pytorch==0.0.4 2>&1 | tail -4
  echo "== baseline/_ref"; ls baseline/_ref 2>&1 | head -3
} > $out 2>&1
cat $out
