for cfg in "1 1920 1080 50" "1 1200 600 0"; do set -- $cfg
  echo "### scene $1 grid $4"
  for i in 1 2; do python tools/prof_cmd.py 128 $1 $2 $3 $4; RT_LIB=$PWD/accelerated-ray-tracer_b200/lib/variants/lbvh.so python tools/prof_cmd.py 128 $1 $2 $3 $4; done
done
