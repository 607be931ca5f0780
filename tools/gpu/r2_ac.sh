B=remhos_b200/host/remhos
M=tests/data/periodic-hexagon.mesh
run() { echo "== $*"; $B -no-vis -m $M "$@" 2>&1 | grep -i "final\|max\|mass" | tail -3; $B -no-vis -m $M "$@" -gpus 2 2>&1 | grep -i "final\|max\|mass" | tail -3; }
run -p 0 -rs 2 -o 3 -dt 0.005 -tf 0.05 -ho 1 -lo 2 -fct 1 -s 2
run -p 0 -rs 2 -o 3 -dt 0.005 -tf 0.05 -ho 1 -lo 2 -fct 2 -s 2
run -p 0 -rs 2 -o 2 -dt 0.005 -tf 0.05 -ho 1 -lo 2 -fct 1 -s 2
run -p 0 -rs 2 -o 3 -dt 0.005 -tf 0.05 -ho 3 -lo 2 -fct 1 -s 2
run -p 0 -rs 2 -o 3 -dt 0.005 -tf 0.05 -ho 1 -lo 1 -fct 1 -s 2
run -p 0 -rs 2 -o 3 -dt 0.005 -tf 0.05 -ho 1 -lo 2 -fct 1 -s 1
run -p 0 -rs 2 -o 3 -dt 0.005 -tf 0.005 -ho 1 -lo 2 -fct 1 -s 1
