import sys; sys.path.insert(0,"tests"); sys.path.insert(0,".")
import numpy as np, golden_util, oracle_lib
from zkevm_specs_b200 import native
from zkevm_specs_b200.evm_circuit import main as evm_main
from zkevm_specs_b200.evm_circuit.table import fixed_table_matrix
ctx=native.default_context(); fixed=fixed_table_matrix(); cat=native.constraint_catalogue(3)
z=np.load("tests/golden/evm.npz")
nbad=0
lo,hi=(int(sys.argv[1]),int(sys.argv[2])) if len(sys.argv)>2 else (0,10**9)
for rep in range(1):
  for name,k,s,b,r,flags,er,ee in golden_util.evm_vectors():
    if name!='add_sub' or not (lo<=k<=hi): continue
    ctx.upload_table(native.TABLE_BYTECODE,b); ctx.upload_table(native.TABLE_RW,r); evm_main.upload_fixed_table(ctx)
    ctx.upload_columns(native.CIRCUIT_EVM,s)
    ff,fc=ctx.check(native.CIRCUIT_EVM,0,s.shape[1]-1,0,flags)
    off,ofc=oracle_lib.check_evm(s,b,r,fixed,flags=flags)
    if not (np.array_equal(ff,off) and np.array_equal(fc,ofc)):
        nbad+=1
        if nbad<12:
            print(rep,name,k,"kind",int(z[f"{name}/mut_kind"][k]),"row",int(z[f"{name}/mut_row"][k]),"col",int(z[f"{name}/mut_col"][k]))
            print("   gpu",[(cat[i][0][:26],int(ff[i]),int(fc[i])) for i in np.nonzero((ff!=0xFFFFFFFF)|(fc!=0))[0]])
            print("   orc",[(cat[i][0][:26],int(off[i]),int(ofc[i])) for i in np.nonzero((off!=0xFFFFFFFF)|(ofc!=0))[0]])
print("mismatches",nbad)
