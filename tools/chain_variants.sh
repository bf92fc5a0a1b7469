# times the compiled scheduling variants of the chain kernel (GVN_TC_VARIANT) at the benchmark shape
for v in ${@:-84 86 118 22 84}; do echo "variant $v"; GVN_TC_VARIANT=$v python tools/estep_time.py; done
