# compute-sanitizer on the kernel families changed in this session (mstft: merged launches, cp.async + mbarrier table fill; Griffin-Lim:
# packed previous spectrum)
O=gpurun_out/r02c/san; mkdir -p $O
for t in memcheck racecheck synccheck initcheck; do
  SB200_SANITIZE_ONLY=mstft,stft_complex timeout 600 compute-sanitizer --tool $t --target-processes all --error-exitcode 1 python tools/sanitize_cases.py > $O/san_${t}_mstft.log 2>&1; echo "$t mstft rc=$? $(grep -c "ERROR SUMMARY\|RACECHECK SUMMARY" $O/san_${t}_mstft.log) $(grep "SUMMARY" $O/san_${t}_mstft.log | tail -1)"
done
for t in memcheck racecheck initcheck; do
  SB200_SANITIZE_ONLY=gl,gl_multilaunch timeout 600 compute-sanitizer --tool $t --target-processes all --error-exitcode 1 python tools/sanitize_cases.py > $O/san_${t}_gl.log 2>&1; echo "$t gl rc=$? $(grep "SUMMARY" $O/san_${t}_gl.log | tail -1)"
done
for f in $O/*.log; do head -c 20000 $f > $f.tmp && mv $f.tmp $f; done
