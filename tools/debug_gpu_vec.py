import sys; sys.path.insert(0,"tests"); sys.path.insert(0,".")
import numpy as np, golden_util, oracle_lib
from zkevm_specs_b200 import native
from zkevm_specs_b200.evm_circuit import main as evm_main
from zkevm_specs_b200.evm_circuit.table import fixed_table_matrix
ctx=native.default_context(); fixed=fixed_table_matrix(); cat=native.constraint_catalogue(3)
want=set(int(x) for x in sys.argv[1:]) or {71}
for name,k,s,b,r,flags,er,ee in golden_util.evm_vectors():
    if name!="add_sub" or k not in want: continue
    for rep in range(3):
        ctx.upload_table(native.TABLE_BYTECODE,b); ctx.upload_table(native.TABLE_RW,r); evm_main.upload_fixed_table(ctx)
        ctx.upload_columns(native.CIRCUIT_EVM,s)
        ff,fc=ctx.check(native.CIRCUIT_EVM,0,s.shape[1]-1,0,flags)
        print(k,"gpu",[(cat[i][0][:28],int(ff[i]),int(fc[i])) for i in np.nonzero(ff!=0xFFFFFFFF)[0]])
    off,ofc=oracle_lib.check_evm(s,b,r,fixed,flags=flags)
    print(k,"orc",[(cat[i][0][:28],int(off[i]),int(ofc[i])) for i in np.nonzero(off!=0xFFFFFFFF)[0]])
