for pass in 1 2; do
  (cd build/r1_tree && python tools/ab.py --rounds 5 --steps 400 --modes step manytor_b200/lib/libmanytor_b200.so) 2>&1 | grep median | sed 's/^/r1   /'
  python tools/ab.py --rounds 5 --steps 400 --modes step build/variants/r2base.so 2>&1 | grep median | sed 's/^/base /'
  python tools/ab.py --rounds 5 --steps 400 --modes step manytor_b200/lib/libmanytor_b200.so 2>&1 | grep median | sed 's/^/head /'
done
for pass in 1 2; do
  (cd build/r1_tree && python tools/ab.py --rounds 5 --steps 300 --modes step --arm ur5 --x 20 manytor_b200/lib/libmanytor_b200.so) 2>&1 | grep median | sed 's/^/r1   ur5 /'
  python tools/ab.py --rounds 5 --steps 300 --modes step --arm ur5 --x 20 build/variants/r2base.so 2>&1 | grep median | sed 's/^/base ur5 /'
  python tools/ab.py --rounds 5 --steps 300 --modes step --arm ur5 --x 20 manytor_b200/lib/libmanytor_b200.so 2>&1 | grep median | sed 's/^/head ur5 /'
done
