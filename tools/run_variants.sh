# usage: bash tools/run_variants.sh "<variant names, '' = product>" <command...>   (tuning variants: see build.py, TITGPU_VARIANT)
names="$1"; shift
for v in $names; do
  if [ "$v" != base ]; then export TITGPU_LIB=$PWD/titsolver_b200/libtitgpu_$v.so; else unset TITGPU_LIB; fi
  echo "== $v"; "$@"
done
