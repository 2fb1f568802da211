for v in main m16nosplit m16mb6 mini8 m8nosplit; do
  if [ $v = main ]; then unset ICER_B200_LIB; else export ICER_B200_LIB=$PWD/icer_compression_b200/lib/variants/libicer_b200_$v.so; fi
  echo "== $v"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
done
bash profiles/run_variants2.sh m16nosplit m16mb6 mini8 m8nosplit
