"""Host-side witness builders with the reference's names: `Bytecode` (opcode DSL),
`RWDictionary`, `KeccakCircuit`, `CopyCircuit`, `Block`, `Transaction`, `Withdrawal`, `Account`.

Mirrors /root/reference/src/zkevm_specs/evm_circuit/typing.py:327-427 (Bytecode), :464-845
(RWDictionary), :848-865 (KeccakCircuit), :996-1150 (CopyCircuit).  They stay Python (SURVEY.md
§8 a20): they produce the rows that packing.py turns into cell matrices for the device."""
from __future__ import annotations

import dataclasses
from typing import Dict, Iterator, List, Mapping, MutableSequence, NamedTuple, Optional, Sequence, Tuple, Union

from ..util.arithmetic import FQ, RLC, IntOrFQ, Word, WordOrValue
from ..util.hash import keccak256
from .spec import (RW, AccountFieldTag, BlockContextFieldTag, BytecodeFieldTag, CallContextFieldTag,
                   CopyDataTypeTag, Opcode, Target, TxContextFieldTag, TxLogFieldTag, TxReceiptFieldTag, get_push_size)
from .table import (BlockTableRow, BytecodeTableRow, CopyCircuitRow, KeccakTableRow, RWTableRow, TxTableRow,
                    WithdrawalTableRow)


def init_is_code(code: bytes) -> List[bool]:
    flags, left = [], 0
    for b in code:
        flags.append(left == 0)
        left = get_push_size(b) if left == 0 else left - 1
    return flags


class Bytecode:
    """Bytecode builder: `Bytecode().push32(1).add().stop()` (reference typing.py:327-427)."""

    def __init__(self, code: Optional[bytearray] = None, is_code: Optional[MutableSequence[bool]] = None):
        self.code = bytearray() if code is None else bytearray(code)
        self.is_code = init_is_code(self.code) if is_code is None else list(is_code)

    def __getattr__(self, name: str):
        if name.startswith("__"):
            raise AttributeError(name)
        key = name[:-1].upper() if name.endswith("_") else name.upper()
        if key not in Opcode.__members__:
            raise ValueError(f"Invalid opcode {name}")
        opcode = Opcode[key]

        def emit(*args) -> "Bytecode":
            if opcode.is_push_with_data():
                assert len(args) == 1
                return self.push(args[0], int(opcode) - int(Opcode.PUSH0))
            if not (opcode.is_dup() or opcode.is_swap()):
                assert len(args) <= 1024 - opcode.max_stack_pointer()
                for arg in reversed(args):
                    self.push(arg)
            else:
                assert len(args) == 0
            self.code.append(int(opcode))
            self.is_code.append(True)
            return self

        return emit

    def push(self, value, n_bytes: int = 32) -> "Bytecode":
        if isinstance(value, Word):
            value = value.int_value().to_bytes(n_bytes, "big")
        elif isinstance(value, RLC):
            value = bytes(reversed(value.le_bytes))
        elif isinstance(value, FQ):
            value = value.n.to_bytes(n_bytes, "big")
        elif isinstance(value, int):
            value = int(value).to_bytes(n_bytes, "big")
        elif isinstance(value, str):
            value = bytes.fromhex(value.lower().removeprefix("0x"))
        elif not isinstance(value, (bytes, bytearray)):
            raise NotImplementedError(f"Value of type {type(value)} is not yet supported")
        assert len(value) <= n_bytes, ValueError("Too many bytes as data portion of PUSH*")
        self.code.append(int(Opcode.PUSH0) + n_bytes)
        self.is_code.append(True)
        self.code.extend(bytes(value).rjust(n_bytes, b"\x00"))
        self.is_code.extend([False] * n_bytes)
        return self

    def hash(self) -> int:
        return int.from_bytes(keccak256(bytes(self.code)), "big")

    def table_assignments(self) -> Iterator[BytecodeTableRow]:
        """Header row (value = length) then one Byte row per byte (typing.py:390-427)."""
        h = Word(self.hash())
        yield BytecodeTableRow(h, FQ(BytecodeFieldTag.Header), FQ(0), FQ(0), FQ(len(self.code)))
        for idx, (byte, is_code) in enumerate(zip(self.code, self.is_code)):
            yield BytecodeTableRow(h, FQ(BytecodeFieldTag.Byte), FQ(idx), FQ(int(is_code)), FQ(byte))


class Block:
    """Block context rows (reference typing.py:70-136); the hot-path gadgets never read them,
    the class exists so that `Tables(block_table=set(Block().table_assignments()), ...)` reads
    like the reference's tests."""

    def __init__(self, coinbase: int = 0x10, gas_limit: int = int(15e6), number: int = 0,
                 timestamp: int = 0, prev_randao: int = 0, base_fee: int = int(1e9), chainid: int = 0x01,
                 withdrawal_root: int = 0, history_hashes: Sequence[int] = ()) -> None:
        assert len(history_hashes) <= min(256, number)
        self.coinbase, self.gas_limit, self.number, self.timestamp = coinbase, gas_limit, number, timestamp
        self.prev_randao, self.base_fee, self.chainid = prev_randao, base_fee, chainid
        self.withdrawal_root = withdrawal_root
        self.history_hashes = list(history_hashes)

    def table_assignments(self) -> List[BlockTableRow]:
        T = BlockContextFieldTag
        fields = [(T.Coinbase, self.coinbase, True), (T.GasLimit, self.gas_limit, False),
                  (T.Number, self.number, False), (T.Timestamp, self.timestamp, False),
                  (T.PrevRandao, self.prev_randao, True), (T.BaseFee, self.base_fee, True),
                  (T.ChainId, self.chainid, False), (T.WithdrawalRoot, self.withdrawal_root, False)]
        rows = [BlockTableRow(FQ(t), FQ(0), WordOrValue(Word(v) if is_word else FQ(v))) for t, v, is_word in fields]
        for back, h in enumerate(reversed(self.history_hashes)):
            rows.append(BlockTableRow(FQ(T.HistoryHash), FQ(self.number - back - 1), WordOrValue(Word(h))))
        return rows


class AccessTuple(NamedTuple):
    address: int
    storage_keys: List[int]


# gas schedule of a transaction's payload (reference util/param.py: call data per zero / non-zero byte,
# access list per address / storage key)
_GAS_CALLDATA_ZERO, _GAS_CALLDATA_NONZERO, _GAS_AL_ADDRESS, _GAS_AL_STORAGE = 4, 16, 2400, 1900


class Transaction:
    """One transaction and its tx-table rows (reference typing.py:145-281): twelve fixed fields per tx id,
    then one CallData row per byte."""

    def __init__(self, id: int = 1, nonce: int = 0, gas: int = 21000, gas_price: int = int(2e9), caller_address: int = 0xCAFE,
                 callee_address: Optional[int] = None, value: int = 0, call_data: bytes = bytes(), invalid_tx: int = 0,
                 access_list: Sequence[AccessTuple] = ()) -> None:
        self.id, self.nonce, self.gas, self.gas_price = id, nonce, gas, gas_price
        self.caller_address, self.callee_address, self.value = caller_address, callee_address, value
        self.call_data, self.invalid_tx, self.access_list = call_data, invalid_tx, list(access_list)

    @classmethod
    def padding(cls, id: int) -> "Transaction":
        return cls(id, 0, 0, 0, 0, 0, 0, bytes(), 0, [])

    def call_data_gas_cost(self) -> int:
        zeros = sum(1 for b in self.call_data if b == 0)
        return zeros * _GAS_CALLDATA_ZERO + (len(self.call_data) - zeros) * _GAS_CALLDATA_NONZERO

    def access_list_gas_cost(self) -> int:
        return sum(_GAS_AL_ADDRESS + len(t.storage_keys) * _GAS_AL_STORAGE for t in self.access_list)

    def table_fixed(self) -> List[TxTableRow]:
        T = TxContextFieldTag
        fields = [(T.Nonce, self.nonce, False), (T.Gas, self.gas, False), (T.GasPrice, self.gas_price, True),
                  (T.CallerAddress, self.caller_address, True),
                  (T.CalleeAddress, 0 if self.callee_address is None else self.callee_address, True),
                  (T.IsCreate, int(self.callee_address is None), False), (T.Value, self.value, True),
                  (T.CallDataLength, len(self.call_data), False), (T.CallDataGasCost, self.call_data_gas_cost(), False),
                  (T.TxInvalid, self.invalid_tx, False), (T.AccessListGasCost, self.access_list_gas_cost(), False),
                  (T.TxSignHash, 1234, False)]  # the reference's mock sign hash
        return [TxTableRow(FQ(self.id), FQ(t), FQ(0), WordOrValue(Word(v) if is_word else FQ(v))) for t, v, is_word in fields]

    def table_assignments(self) -> Iterator[TxTableRow]:
        yield from self.table_fixed()
        for idx, byte in enumerate(self.call_data):
            yield TxTableRow(FQ(self.id), FQ(TxContextFieldTag.CallData), FQ(idx), WordOrValue(FQ(byte)))


class Withdrawal:
    """One validator withdrawal and its table row (reference typing.py:284-316); amount 0 = padding"""

    def __init__(self, id: int = 0, validator_id: int = 0, address: int = 0xCAFE, amount: int = int(1e9)) -> None:
        self.id, self.validator_id, self.address, self.amount = id, validator_id, address, amount

    @classmethod
    def padding(cls, id: int) -> "Withdrawal":
        return cls(id, 0, 0, 0)

    def table_assignments(self) -> List[WithdrawalTableRow]:
        return [WithdrawalTableRow(FQ(self.id), FQ(self.validator_id), FQ(self.address), FQ(self.amount))]


class Account:
    """An account as the tests describe one (reference typing.py:433-461)"""

    def __init__(self, address: int = 0, nonce: int = 0, balance: int = 0, code: Optional["Bytecode"] = None,
                 storage: Optional[Dict[int, int]] = None) -> None:
        self.address, self.nonce, self.balance = address, nonce, balance
        self.code = Bytecode() if code is None else code
        self.storage = dict() if storage is None else storage

    def code_hash(self) -> int:
        return self.code.hash()

    def is_empty(self) -> bool:
        from ..util.hash import EMPTY_CODE_HASH
        return self.nonce == 0 and self.balance == 0 and self.code_hash() == EMPTY_CODE_HASH


_WORD_CALL_CONTEXT = ("CallerAddress", "CalleeAddress", "Value", "CodeHash")


class RWDictionary:
    """Sequential RW-table builder: each call appends one row and bumps rw_counter
    (reference typing.py:464-845)."""

    def __init__(self, rw_counter: int) -> None:
        self.rw_counter = rw_counter
        self.rws: List[RWTableRow] = []

    def _append(self, rw, target, id=FQ(0), address=FQ(0), field_tag=FQ(0), storage_key=None,
                value=FQ(0), value_prev=FQ(0), aux0=None, rw_counter: Optional[int] = None) -> "RWDictionary":
        """one row; `rw_counter` given = a row placed out of sequence (the reversion copy of a state write),
        which does not advance the dictionary's counter"""
        as_fq = lambda v: FQ(v) if isinstance(v, int) else v  # noqa: E731
        at = self.rw_counter if rw_counter is None else rw_counter
        self.rws.append(RWTableRow(
            FQ(at), FQ(rw), FQ(target), as_fq(id), as_fq(address), as_fq(field_tag),
            Word(0) if storage_key is None else storage_key,
            WordOrValue(as_fq(value)), WordOrValue(as_fq(value_prev)),
            Word(0) if aux0 is None else aux0))
        if rw_counter is None:
            self.rw_counter += 1
        return self

    def _state_write(self, target, rw_counter_of_reversion: Optional[int] = None, **cells) -> "RWDictionary":
        """a reversible write (Target.write_with_reversion, reference typing.py:750-788): the write itself and,
        when the call will revert, its undo row (value and value_prev swapped) at rw_counter_of_reversion"""
        self._append(RW.Write, target, **cells)
        if rw_counter_of_reversion is not None:
            undo = dict(cells)
            undo["value"], undo["value_prev"] = cells.get("value_prev", FQ(0)), cells.get("value", FQ(0))
            self._append(RW.Write, target, rw_counter=rw_counter_of_reversion, **undo)
        return self

    def stack_read(self, call_id: IntOrFQ, stack_pointer: IntOrFQ, value: Word) -> "RWDictionary":
        return self._append(RW.Read, Target.Stack, id=FQ(call_id), address=FQ(stack_pointer), value=value)

    def stack_write(self, call_id: IntOrFQ, stack_pointer: IntOrFQ, value: Word) -> "RWDictionary":
        return self._append(RW.Write, Target.Stack, id=FQ(call_id), address=FQ(stack_pointer), value=value)

    def memory_read(self, call_id: IntOrFQ, memory_address: IntOrFQ, byte: IntOrFQ) -> "RWDictionary":
        return self._append(RW.Read, Target.Memory, id=FQ(call_id), address=FQ(memory_address), value=FQ(byte))

    def memory_write(self, call_id: IntOrFQ, memory_address: IntOrFQ, byte: IntOrFQ) -> "RWDictionary":
        return self._append(RW.Write, Target.Memory, id=FQ(call_id), address=FQ(memory_address), value=FQ(byte))

    def _call_context(self, rw, call_id, field_tag, value) -> "RWDictionary":
        if isinstance(value, int):
            value = FQ(value)
        if CallContextFieldTag(field_tag).name in _WORD_CALL_CONTEXT:
            assert isinstance(value, Word)
        else:
            assert isinstance(value, FQ)
        return self._append(rw, Target.CallContext, id=FQ(call_id), address=FQ(field_tag), value=value)

    def call_context_read(self, call_id: IntOrFQ, field_tag, value) -> "RWDictionary":
        return self._call_context(RW.Read, call_id, field_tag, value)

    def call_context_write(self, call_id: IntOrFQ, field_tag, value) -> "RWDictionary":
        return self._call_context(RW.Write, call_id, field_tag, value)

    def tx_log_write(self, tx_id: IntOrFQ, log_id: int, field_tag, index: IntOrFQ, value) -> "RWDictionary":
        if isinstance(value, int):
            value = FQ(value)
        if TxLogFieldTag(field_tag) in (TxLogFieldTag.Address, TxLogFieldTag.Topic):
            assert isinstance(value, Word)
        else:
            assert isinstance(value, FQ)
        return self._append(RW.Write, Target.TxLog, id=FQ(tx_id),
                            address=FQ(int(index) + (int(field_tag) << 32) + (log_id << 48)), value=value)

    def tx_refund_read(self, tx_id: IntOrFQ, refund: IntOrFQ) -> "RWDictionary":
        return self._append(RW.Read, Target.TxRefund, id=FQ(tx_id), value=FQ(refund), value_prev=FQ(refund))

    def tx_refund_write(self, tx_id: IntOrFQ, refund: IntOrFQ, refund_prev: IntOrFQ,
                        rw_counter_of_reversion: Optional[int] = None) -> "RWDictionary":
        return self._state_write(Target.TxRefund, rw_counter_of_reversion, id=FQ(tx_id), value=FQ(refund), value_prev=FQ(refund_prev))

    def tx_receipt_read(self, tx_id: IntOrFQ, field_tag, value: IntOrFQ) -> "RWDictionary":
        return self._append(RW.Read, Target.TxReceipt, id=FQ(tx_id), field_tag=FQ(field_tag), value=FQ(value))

    def tx_receipt_write(self, tx_id: IntOrFQ, field_tag, value: IntOrFQ) -> "RWDictionary":
        return self._append(RW.Write, Target.TxReceipt, id=FQ(tx_id), field_tag=FQ(field_tag), value=FQ(value))

    def tx_access_list_account_write(self, tx_id: IntOrFQ, account_address: IntOrFQ, value: bool, value_prev: bool,
                                     rw_counter_of_reversion: Optional[int] = None) -> "RWDictionary":
        return self._state_write(Target.TxAccessListAccount, rw_counter_of_reversion, id=FQ(tx_id), address=FQ(account_address),
                                 value=FQ(value), value_prev=FQ(value_prev))

    def tx_access_list_account_read(self, tx_id: IntOrFQ, account_address: IntOrFQ, value: bool) -> "RWDictionary":
        return self._append(RW.Read, Target.TxAccessListAccount, id=FQ(tx_id), address=FQ(account_address),
                            value=FQ(value), value_prev=FQ(value))

    def tx_access_list_account_storage_write(self, tx_id: IntOrFQ, account_address: IntOrFQ, storage_key: Word, value: bool,
                                             value_prev: bool, rw_counter_of_reversion: Optional[int] = None) -> "RWDictionary":
        return self._state_write(Target.TxAccessListAccountStorage, rw_counter_of_reversion, id=FQ(tx_id),
                                 address=FQ(account_address), storage_key=storage_key, value=FQ(value), value_prev=FQ(value_prev))

    def tx_access_list_account_storage_read(self, tx_id: IntOrFQ, account_address: IntOrFQ, storage_key: Word,
                                            value: bool) -> "RWDictionary":
        return self._append(RW.Read, Target.TxAccessListAccountStorage, id=FQ(tx_id), address=FQ(account_address),
                            storage_key=storage_key, value=FQ(value), value_prev=FQ(value))

    def account_read(self, account_address: IntOrFQ, field_tag, value) -> "RWDictionary":
        if isinstance(value, int):
            value = FQ(value)
        return self._append(RW.Read, Target.Account, address=FQ(account_address), field_tag=FQ(field_tag),
                            value=value, value_prev=value)

    def account_write(self, account_address: IntOrFQ, field_tag, value, value_prev,
                      rw_counter_of_reversion: Optional[int] = None) -> "RWDictionary":
        return self._state_write(Target.Account, rw_counter_of_reversion, address=FQ(account_address), field_tag=FQ(field_tag),
                                 value=value, value_prev=value_prev)

    def account_storage_read(self, account_address: IntOrFQ, storage_key: Word, value: Word, tx_id: IntOrFQ,
                             value_committed: Word) -> "RWDictionary":
        return self._append(RW.Read, Target.AccountStorage, id=tx_id, address=FQ(account_address), storage_key=storage_key,
                            value=value, value_prev=value, aux0=value_committed)

    def account_storage_write(self, account_address: IntOrFQ, storage_key: Word, value: Word, value_prev: Word, tx_id: IntOrFQ,
                              value_committed: Word, rw_counter_of_reversion: Optional[int] = None) -> "RWDictionary":
        return self._state_write(Target.AccountStorage, rw_counter_of_reversion, id=tx_id, address=FQ(account_address),
                                 storage_key=storage_key, value=value, value_prev=value_prev, aux0=value_committed)


class KeccakCircuit:
    """Keccak table rows (state_tag=2 Finalize, input_rlc, input_len, output) — reference
    typing.py:848-865.  Hashing is witness generation and stays on the host."""

    def __init__(self) -> None:
        self.rows: List[KeccakTableRow] = []

    def add(self, data: bytes, r: FQ) -> "KeccakCircuit":
        output = Word(int.from_bytes(keccak256(bytes(data)), "big"))
        acc = RLC(bytes(reversed(bytes(data))), r, n_bytes=len(data))
        self.rows.append(KeccakTableRow(FQ(2), acc.expr(), FQ(len(data)), output))
        return self


class CopyCircuit:
    """Copy-circuit witness: two rows (read, write) per copied byte (reference
    typing.py:996-1150).  `copy()` also appends the memory / tx-log rows it touches to the
    RWDictionary, like the reference."""

    def __init__(self, pad_rows: Optional[List[CopyCircuitRow]] = None) -> None:
        self.rows: List[CopyCircuitRow] = []
        self.pad_rows: List[CopyCircuitRow] = list(pad_rows) if pad_rows is not None else []

    def table(self) -> Sequence[CopyCircuitRow]:
        return self.rows + self.pad_rows

    def copy(self, r: FQ, rw_dict: RWDictionary, src_id, src_tag, dst_id, dst_tag, src_addr: IntOrFQ,
             src_addr_end: IntOrFQ, dst_addr: IntOrFQ, copy_length: IntOrFQ,
             src_data: Mapping, log_id: int = 0) -> "CopyCircuit":
        n = int(copy_length)
        src_addr, src_addr_end, dst_addr = int(src_addr), int(src_addr_end), int(dst_addr)
        uses_code = CopyDataTypeTag.Bytecode in (src_tag, dst_tag)
        new_rows: List[CopyCircuitRow] = []
        acc = FQ(0)
        for i in range(n):
            is_pad = src_addr + i >= src_addr_end
            value, is_code = FQ(0), FQ(0)
            if not is_pad:
                assert src_addr + i in src_data, f"Cannot find data at the offset {src_addr + i}"
                item = src_data[src_addr + i]
                if uses_code:
                    value, is_code = FQ(item[0]), FQ(item[1])
                else:
                    value = FQ(item)
            self._emit(new_rows, rw_dict, False, i == 0, False, src_id, src_tag, src_addr + i, value,
                       is_code, is_pad, src_addr_end=src_addr_end, bytes_left=n - i)
            if dst_tag == CopyDataTypeTag.RlcAcc:
                acc = acc * r + value
            self._emit(new_rows, rw_dict, True, False, i == n - 1, dst_id, dst_tag, dst_addr + i,
                       acc if dst_tag == CopyDataTypeTag.RlcAcc else value, is_code, False, log_id=log_id)
        end = rw_dict.rw_counter
        for row in new_rows:
            upd = {"rwc_inc_left": FQ(end - row.rw_counter.n)}
            if dst_tag == CopyDataTypeTag.RlcAcc:
                upd["rlc_acc"] = acc
            self.rows.append(dataclasses.replace(row, **upd))
        return self

    @staticmethod
    def _emit(rows, rw_dict: RWDictionary, is_write: bool, is_first: bool, is_last: bool, id, tag,
              addr: int, value: FQ, is_code: FQ, is_pad: bool, src_addr_end: int = 0,
              bytes_left: int = 0, log_id: int = 0) -> None:
        id_cell = WordOrValue(id if isinstance(id, Word) else FQ(id))
        rw_counter = rw_dict.rw_counter
        if tag == CopyDataTypeTag.Memory and not is_pad:
            (rw_dict.memory_write if is_write else rw_dict.memory_read)(id_cell.value(), addr, value)
        elif tag == CopyDataTypeTag.TxLog:
            assert is_write
            rw_dict.tx_log_write(id_cell.value(), log_id, TxLogFieldTag.Data, addr, value)
            addr += (int(TxLogFieldTag.Data) << 32) + (log_id << 48)
        T = CopyDataTypeTag
        rows.append(CopyCircuitRow(
            q_step=FQ(int(not is_write)), is_first=FQ(int(is_first)), is_last=FQ(int(is_last)),
            id=id_cell, tag=FQ(tag), addr=FQ(addr), src_addr_end=FQ(src_addr_end),
            bytes_left=FQ(bytes_left), value=value, rlc_acc=FQ(0), is_code=is_code,
            is_pad=FQ(int(is_pad)), rw_counter=FQ(rw_counter), rwc_inc_left=FQ(0),
            is_memory=FQ(int(tag == T.Memory)), is_bytecode=FQ(int(tag == T.Bytecode)),
            is_tx_calldata=FQ(int(tag == T.TxCalldata)), is_tx_log=FQ(int(tag == T.TxLog)),
            is_rlc_acc=FQ(int(tag == T.RlcAcc))))
