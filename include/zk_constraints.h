/*
 * zk_constraints.h — constraint catalogue: the ids that index first_fail[] / fail_count[].
 *
 * Ids are numbered in the reference's PROGRAM ORDER inside one row / step, so that
 * "the exception the reference raises first" == the failing id with the smallest
 * (row, id) pair.  X(id_name, error_class, "reference file:line — what it asserts").
 * Both the CUDA product and the CPU oracle include this file; it contains no code.
 */
#ifndef ZK_CONSTRAINTS_H
#define ZK_CONSTRAINTS_H

/* error classes mirror ZK_ERR_* in zkcheck.h */
#define ZKE_ASSERT 0
#define ZKE_UNSAT 1
#define ZKE_AMBIG 2
#define ZKE_RANGE 3
#define ZKE_VALUE 4
#define ZKE_NOTIMPL 5

/* ---------------- bytecode circuit: src/zkevm_specs/bytecode_circuit.py:37-100 ------- */
#define ZK_BYTECODE_CONSTRAINTS(X)                                                        \
  X(BC_FIRST_TAG, ZKE_ASSERT, "bytecode_circuit.py:44-45 q_first => tag==Header")         \
  X(BC_HDR_VALUE_LEN, ZKE_ASSERT, "bytecode_circuit.py:49 header value==length")          \
  X(BC_HDR_INDEX0, ZKE_ASSERT, "bytecode_circuit.py:50 header index==0")                  \
  X(BC_H2B_LEN, ZKE_ASSERT, "bytecode_circuit.py:73 next.length==length")                 \
  X(BC_H2B_INDEX0, ZKE_ASSERT, "bytecode_circuit.py:74 next.index==0")                    \
  X(BC_H2B_ISCODE, ZKE_ASSERT, "bytecode_circuit.py:75 next.is_code==1")                  \
  X(BC_H2B_HASH, ZKE_ASSERT, "bytecode_circuit.py:76 next.hash==hash")                    \
  X(BC_H2B_RLC, ZKE_ASSERT, "bytecode_circuit.py:77 next.value_rlc==next.value")          \
  X(BC_H2H_LEN0, ZKE_ASSERT, "bytecode_circuit.py:81 length==0")                          \
  X(BC_H2H_EMPTY_HASH, ZKE_ASSERT, "bytecode_circuit.py:82 hash==EMPTY_HASH")             \
  X(BC_PUSH_TABLE, ZKE_ASSERT, "bytecode_circuit.py:57 (value,push_data_size) in push_table") \
  X(BC_IS_CODE, ZKE_ASSERT, "bytecode_circuit.py:58 is_code==(push_data_left==0)")        \
  X(BC_B2B_LEN, ZKE_ASSERT, "bytecode_circuit.py:87 next.length==length")                 \
  X(BC_B2B_INDEX, ZKE_ASSERT, "bytecode_circuit.py:88 next.index==index+1")               \
  X(BC_B2B_HASH, ZKE_ASSERT, "bytecode_circuit.py:89 next.hash==hash")                    \
  X(BC_B2B_RLC, ZKE_ASSERT, "bytecode_circuit.py:90 next.value_rlc==value_rlc*r+next.value") \
  X(BC_B2B_PUSH_LEFT, ZKE_ASSERT, "bytecode_circuit.py:91-94 push_data_left transition")  \
  X(BC_B2H_INDEX, ZKE_ASSERT, "bytecode_circuit.py:99 index+1==length")                   \
  X(BC_B2H_KECCAK, ZKE_ASSERT, "bytecode_circuit.py:100 (2,value_rlc,length,hash) in keccak_table") \
  X(BC_LAST_TAG, ZKE_ASSERT, "bytecode_circuit.py:66 q_last => tag==Header")              \
  X(BC_LAST_LEN0, ZKE_ASSERT, "bytecode_circuit.py:67,81 q_last => length==0")            \
  X(BC_LAST_EMPTY_HASH, ZKE_ASSERT, "bytecode_circuit.py:67,82 q_last => hash==EMPTY_HASH")

/* generic enum builder */
#define ZK_ENUM_ENTRY(id, cls, doc) id,

enum zk_bytecode_constraint { ZK_BYTECODE_CONSTRAINTS(ZK_ENUM_ENTRY) BC_N_CONSTRAINTS };

#endif /* ZK_CONSTRAINTS_H */
