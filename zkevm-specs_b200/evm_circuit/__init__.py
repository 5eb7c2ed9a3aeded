from .spec import *  # noqa: F401,F403
from .spec import (Opcode, ExecutionState, FixedTableTag, BlockContextFieldTag, TxContextFieldTag,  # noqa: F401
                   BytecodeFieldTag, RW, Target, CallContextFieldTag, AccountFieldTag, TxLogFieldTag,
                   TxReceiptFieldTag, CopyDataTypeTag, MPTProofType, is_push_with_data, get_push_size)
from .table import *  # noqa: F401,F403
from .table import (Tables, FixedTableRow, BlockTableRow, TxTableRow, WithdrawalTableRow, BytecodeTableRow, RWTableRow,  # noqa: F401
                    MPTTableRow, CopyCircuitRow, CopyTableRow, KeccakTableRow, LookupUnsatFailure,
                    LookupAmbiguousFailure)
from .typing import (AccessTuple, Account, Bytecode, Block, RWDictionary, KeccakCircuit, CopyCircuit, Transaction,  # noqa: F401
                     Withdrawal)
