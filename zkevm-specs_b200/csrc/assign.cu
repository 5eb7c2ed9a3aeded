// assign.cu — witness assignment on the device: the row builders the reference runs as
// Python object loops, fed from the compact data they expand.
//
//   assign_bytecode_circuit  src/zkevm_specs/bytecode_circuit.py:104-167   raw code bytes -> 12-cell rows
//   op2row / assign_state_circuit  src/zkevm_specs/state_circuit.py:827-889  15 operation cells -> 57-cell rows
//   CopyCircuit.copy         src/zkevm_specs/evm_circuit/typing.py:1010-1147 copy events + bytes -> 20-cell rows
//
// The two sequential pieces — value_rlc of a contract and rlc_acc of a copy event, both Horner chains
// acc = acc * r + byte — run as a chunked segmented scan: one thread folds a 32-byte chunk
// (k_seg_local), one thread per segment chains the chunk values with r^32 (k_seg_carry), then one
// thread per chunk replays its bytes from the carried-in value and writes the running value of each row.
// Everything else of a row is a closed form of (segment, position).  Rows are written as narrow columns
// (fr.cuh:ld_col widths), so a check after an assignment reads 80-150 B per row instead of 384-1,824.
// Included by api.cu (it uses the context).
#include "circuit.cuh"

namespace zk {

#define ZK_SEG_CHUNK 32
struct SegHorner {
  const unsigned char* bytes;  // the segments' bytes, concatenated
  const u64* seg_off;          // [n_seg + 1] byte offsets
  const u64* chunk_off;        // [n_seg + 1] first chunk of each segment (ceil(len / 32) chunks each)
  u64 n_seg, n_chunks;
  Fr r_mont;  // the challenge in Montgomery form
};

ZK_HD u64 seg_of_chunk(const SegHorner& s, u64 c) {  // the largest k with chunk_off[k] <= c
  u64 lo = 0, hi = s.n_seg;
  while (hi - lo > 1) {
    const u64 mid = (lo + hi) >> 1;
    if (s.chunk_off[mid] <= c) lo = mid;
    else hi = mid;
  }
  return lo;
}
ZK_HD void st_narrow(unsigned char* col, u32 width, u64 row, u64 v) {
  if (width == 1) col[row] = (unsigned char)v;
  else if (width == 2) ((unsigned short*)col)[row] = (unsigned short)v;
  else if (width == 4) ((u32*)col)[row] = (u32)v;
  else ((u64*)col)[row] = v;
}
ZK_HD void st_u128(unsigned char* col, u64 row, u64 lo, u64 hi) {
  ((u64*)col)[2 * row] = lo;
  ((u64*)col)[2 * row + 1] = hi;
}
ZK_HD void st_fr(unsigned char* col, u64 row, const Fr& v) {
  u64* p = (u64*)col + 4 * row;
  p[0] = v.l[0];
  p[1] = v.l[1];
  p[2] = v.l[2];
  p[3] = v.l[3];
}

ZK_HD u64 zk_min64(u64 a, u64 b) { return a < b ? a : b; }
// the per-thread bodies are __host__ __device__ so that tests/emu can run them serially on the CPU
ZK_HD void seg_local(const SegHorner& s, Fr* chunk_val, u64 c) {
  const u64 k = seg_of_chunk(s, c);
  const u64 b0 = s.seg_off[k] + (c - s.chunk_off[k]) * ZK_SEG_CHUNK;
  const u64 b1 = zk_min64(b0 + ZK_SEG_CHUNK, s.seg_off[k + 1]);
  Fr acc = fr_u64(0);
  for (u64 j = b0; j < b1; j++) acc = fr_add(fr_montmul(acc, s.r_mont), fr_u64(s.bytes[j]));
  chunk_val[c] = acc;
}
// chunk_val[c] becomes the value carried INTO chunk c; seg_total[k] the Horner value of the whole segment
ZK_HD void seg_carry(const SegHorner& s, Fr* chunk_val, Fr* seg_total, u64 k) {
  Fr rp = s.r_mont;  // r^32, Montgomery form
  for (int q = 0; q < 5; q++) rp = fr_montmul(rp, rp);
  const u64 c0 = s.chunk_off[k], c1 = s.chunk_off[k + 1], len = s.seg_off[k + 1] - s.seg_off[k];
  Fr carry = fr_u64(0);
  for (u64 c = c0; c < c1; c++) {
    const Fr v = chunk_val[c];
    chunk_val[c] = carry;
    const u64 left = len - (c - c0) * ZK_SEG_CHUNK;
    if (left >= ZK_SEG_CHUNK) {
      carry = fr_montmul(carry, rp);
    } else {
      for (u64 q = 0; q < left; q++) carry = fr_montmul(carry, s.r_mont);
    }
    carry = fr_add(carry, v);
  }
  if (seg_total) seg_total[k] = carry;
}

// ---------------------------------------------------------------- bytecode circuit
struct BytecodeAssign {
  SegHorner s;              // segments = contracts
  const unsigned char* bits;  // is_code bit per code byte (Bytecode.table_assignments, typing.py:390-427)
  const u64* hashes;        // [n_seg][4]: code hash lo (2 limbs), hi (2 limbs)
  u64 n_rows;               // 2^k
  u64 n_table_rows;         // total bytes + contracts (what the table would hold), may exceed n_rows
  unsigned char* base;
  u64 off[12];
};
static const unsigned char kBytecodeAssignWidths[12] = {1, 1, 16, 16, 1, 4, 4, 1, 1, 32, 4, 1};
enum { A_QFIRST, A_QLAST, A_HASH_LO, A_HASH_HI, A_TAG, A_INDEX, A_VALUE, A_ISCODE, A_PDL, A_RLC, A_LEN, A_PDS };

// value_rlc of the Byte rows: replay each chunk from its carried-in value
ZK_HD void assign_bytecode_rlc(const BytecodeAssign& a, const Fr* chunk_val, u64 c) {
  const SegHorner& s = a.s;
  const u64 k = seg_of_chunk(s, c);
  const u64 b0 = s.seg_off[k] + (c - s.chunk_off[k]) * ZK_SEG_CHUNK;
  const u64 b1 = zk_min64(b0 + ZK_SEG_CHUNK, s.seg_off[k + 1]);
  Fr acc = chunk_val[c];
  for (u64 j = b0; j < b1; j++) {
    acc = fr_add(fr_montmul(acc, s.r_mont), fr_u64(s.bytes[j]));
    const u64 row = j + k + 1;  // k Header rows of earlier contracts + this one's
    if (row < a.n_rows) st_fr(a.base + a.off[A_RLC], row, acc);
  }
}
// every other cell (and value_rlc = 0 of Header / padding rows): one thread per row
ZK_HD void assign_bytecode_row(const BytecodeAssign& a, u64 r) {
  const SegHorner& s = a.s;
  {
    u64 hl0, hl1, hh0, hh1, tag = 1, index = 0, value = 0, is_code = 0, pdl = 0, len = 0, pds = 0;
    bool byte_row = false;
    if (r < a.n_table_rows) {
      u64 lo = 0, hi = s.n_seg;  // largest k with seg_off[k] + k <= r
      while (hi - lo > 1) {
        const u64 mid = (lo + hi) >> 1;
        if (s.seg_off[mid] + mid <= r) lo = mid;
        else hi = mid;
      }
      const u64 k = lo, start = s.seg_off[k], local = r - (start + k);
      len = s.seg_off[k + 1] - start;
      hl0 = a.hashes[4 * k], hl1 = a.hashes[4 * k + 1], hh0 = a.hashes[4 * k + 2], hh1 = a.hashes[4 * k + 3];
      if (local == 0) {
        value = len;
      } else {
        byte_row = true;
        const u64 j = start + local - 1;
        tag = 2;
        index = local - 1;
        value = s.bytes[j];
        is_code = (a.bits[j >> 3] >> (j & 7)) & 1;
        pds = (value >= 0x60 && value <= 0x7f) ? value - 0x5f : 0;  // get_push_size
        if (!is_code) {  // data byte of the PUSH at the nearest code byte p before it: left = size - (j - p - 1)
          u64 p = j;
          while (p > start) {
            p--;
            if ((a.bits[p >> 3] >> (p & 7)) & 1) break;
          }
          const u64 pv = s.bytes[p];
          const u64 psz = (pv >= 0x60 && pv <= 0x7f) ? pv - 0x5f : 0;
          pdl = psz - (j - p - 1);
        }
      }
    } else {  // padding: (EMPTY_HASH, Header, 0 ...)   bytecode_circuit.py:150-165
      hl0 = 0x7bfad8045d85a470ull, hl1 = 0xe500b653ca82273bull, hh0 = 0x927e7db2dcc703c0ull, hh1 = 0xc5d2460186f7233cull;
    }
    a.base[a.off[A_QFIRST] + r] = r == 0;
    a.base[a.off[A_QLAST] + r] = r == a.n_rows - 1;
    st_u128(a.base + a.off[A_HASH_LO], r, hl0, hl1);
    st_u128(a.base + a.off[A_HASH_HI], r, hh0, hh1);
    a.base[a.off[A_TAG] + r] = (unsigned char)tag;
    ((u32*)(a.base + a.off[A_INDEX]))[r] = (u32)index;
    ((u32*)(a.base + a.off[A_VALUE]))[r] = (u32)value;
    a.base[a.off[A_ISCODE] + r] = (unsigned char)is_code;
    a.base[a.off[A_PDL] + r] = (unsigned char)pdl;
    ((u32*)(a.base + a.off[A_LEN]))[r] = (u32)len;
    a.base[a.off[A_PDS] + r] = (unsigned char)pds;
    if (!byte_row) st_fr(a.base + a.off[A_RLC], r, fr_u64(0));
  }
}

// ---------------------------------------------------------------- state circuit
// op2row: address -> 10 little-endian 16-bit limbs (cells 8..17), storage key -> 32 little-endian bytes
// (cells 18..49); the other 15 cells of a row are the operation's own and stay where they were uploaded.
struct StateAssign {
  const unsigned char* base;  // the uploaded operation columns
  u64 off_addr, off_klo, off_khi;
  unsigned char w_addr, w_klo, w_khi;
  unsigned char* limbs;  // [10] columns of u16
  unsigned char* kbytes;  // [32] columns of u8
  u64 n_rows, limb_stride, byte_stride;
};
ZK_HD void assign_state_row(const StateAssign& a, u64 r) {
  {
    const Fr addr = ld_col(a.base + a.off_addr, a.w_addr, r);
    const Fr klo = ld_col(a.base + a.off_klo, a.w_klo, r), khi = ld_col(a.base + a.off_khi, a.w_khi, r);
#pragma unroll
    for (int q = 0; q < 10; q++)
      ((unsigned short*)(a.limbs + q * a.limb_stride))[r] = (unsigned short)(addr.l[q >> 2] >> (16 * (q & 3)));
#pragma unroll
    for (int q = 0; q < 32; q++) {
      const Fr& h = q < 16 ? klo : khi;
      (a.kbytes + q * a.byte_stride)[r] = (unsigned char)(h.l[(q & 15) >> 3] >> (8 * (q & 7)));
    }
  }
}

// ---------------------------------------------------------------- copy circuit
// one event = CopyCircuit.copy(r, rw_dict, src_id, src_tag, dst_id, dst_tag, src_addr, src_addr_end, dst_addr,
// copy_length, src_data, log_id): 2 rows per byte (read row, write row)
struct CopyEvent {  // 16 u64 per event, host layout of zk_assign_copy_circuit
  u64 src_tag, dst_tag /* low byte: CopyDataTypeTag; bit 8: the id is a Word */, src_addr, src_addr_end, dst_addr, length, log_id, rw_counter;
  u64 src_id[4];  // lo (2 limbs), hi (2 limbs)
  u64 dst_id[4];
};
struct CopyAssign {
  SegHorner s;  // segments = events, bytes = the copied values (0 where the source is out of bounds)
  const CopyEvent* ev;
  const unsigned char* bits;  // is_code bit per byte (events that touch bytecode), may be null
  unsigned char* base;
  unsigned char* flags;  // [n_rows] row flags: bit 0 = the id cell is a Word
  u64 off[20];
  u64 n_rows;
};
static const unsigned char kCopyAssignWidths[20] = {1, 1, 1, 16, 16, 1, 8, 8, 8, 32, 32, 1, 1, 8, 8, 1, 1, 1, 1, 1};
#define CPT_BYTECODE 1
#define CPT_MEMORY 2
#define CPT_TXCALLDATA 3
#define CPT_TXLOG 4
#define CPT_RLCACC 5
#define TXLOG_DATA 3

ZK_HD void assign_copy_chunk(const CopyAssign& a, const Fr* chunk_val, const Fr* seg_total, u64 c) {
  const SegHorner& s = a.s;
  const u64 k = seg_of_chunk(s, c);
  CopyEvent e = a.ev[k];
  const unsigned char src_word = (e.src_tag >> 8) & 1, dst_word = (e.dst_tag >> 8) & 1;  // the id is a Word (row flag bit 0)
  e.src_tag &= 0xFF;
  e.dst_tag &= 0xFF;
  const u64 b0 = s.seg_off[k] + (c - s.chunk_off[k]) * ZK_SEG_CHUNK;
  const u64 b1 = zk_min64(b0 + ZK_SEG_CHUNK, s.seg_off[k + 1]);
  const bool to_rlc = e.dst_tag == CPT_RLCACC, src_mem = e.src_tag == CPT_MEMORY;
  const bool dst_rw = e.dst_tag == CPT_MEMORY || e.dst_tag == CPT_TXLOG;
  const bool has_code = e.src_tag == CPT_BYTECODE || e.dst_tag == CPT_BYTECODE;
  const u64 n_real = e.src_addr_end > e.src_addr ? zk_min64(e.length, e.src_addr_end - e.src_addr) : 0;
  const u64 rw_total = (src_mem ? n_real : 0) + (dst_rw ? e.length : 0);
  const Fr total = to_rlc ? seg_total[k] : fr_u64(0);
  const u64 dst_shift = e.dst_tag == CPT_TXLOG ? ((u64)TXLOG_DATA << 32) + (e.log_id << 48) : 0;
  Fr acc = chunk_val[c];
  unsigned char* B = a.base;
  for (u64 j = b0; j < b1; j++) {
    const u64 i = j - s.seg_off[k];
    const bool pad = i >= n_real;
    const u64 value = pad ? 0 : s.bytes[j];
    const u64 is_code = (has_code && !pad && a.bits) ? (a.bits[j >> 3] >> (j & 7)) & 1 : 0;
    acc = fr_add(fr_montmul(acc, s.r_mont), fr_u64(value));
    const u64 reads_before = src_mem ? zk_min64(i, n_real) : 0, writes_before = dst_rw ? i : 0;
    const u64 rd = 2 * j, wr = 2 * j + 1;
    // read row (typing.py:1043-1058)
    u64 rwc = reads_before + writes_before;
    B[a.off[0] + rd] = 1;
    B[a.off[1] + rd] = i == 0;
    B[a.off[2] + rd] = 0;
    st_u128(B + a.off[3], rd, e.src_id[0], e.src_id[1]);
    st_u128(B + a.off[4], rd, e.src_id[2], e.src_id[3]);
    B[a.off[5] + rd] = (unsigned char)e.src_tag;
    ((u64*)(B + a.off[6]))[rd] = e.src_addr + i;
    ((u64*)(B + a.off[7]))[rd] = e.src_addr_end;
    ((u64*)(B + a.off[8]))[rd] = e.length - i;
    st_fr(B + a.off[9], rd, fr_u64(value));
    st_fr(B + a.off[10], rd, total);
    B[a.off[11] + rd] = (unsigned char)is_code;
    B[a.off[12] + rd] = pad;
    ((u64*)(B + a.off[13]))[rd] = e.rw_counter + rwc;
    ((u64*)(B + a.off[14]))[rd] = rw_total - rwc;
    B[a.off[15] + rd] = e.src_tag == CPT_MEMORY;
    B[a.off[16] + rd] = e.src_tag == CPT_BYTECODE;
    B[a.off[17] + rd] = e.src_tag == CPT_TXCALLDATA;
    B[a.off[18] + rd] = e.src_tag == CPT_TXLOG;
    B[a.off[19] + rd] = e.src_tag == CPT_RLCACC;
    a.flags[rd] = src_word;
    // write row (typing.py:1060-1076)
    rwc = (src_mem ? zk_min64(i + 1, n_real) : 0) + writes_before;
    B[a.off[0] + wr] = 0;
    B[a.off[1] + wr] = 0;
    B[a.off[2] + wr] = i == e.length - 1;
    st_u128(B + a.off[3], wr, e.dst_id[0], e.dst_id[1]);
    st_u128(B + a.off[4], wr, e.dst_id[2], e.dst_id[3]);
    B[a.off[5] + wr] = (unsigned char)e.dst_tag;
    ((u64*)(B + a.off[6]))[wr] = e.dst_addr + i + dst_shift;
    ((u64*)(B + a.off[7]))[wr] = 0;
    ((u64*)(B + a.off[8]))[wr] = 0;
    st_fr(B + a.off[9], wr, to_rlc ? acc : fr_u64(value));
    st_fr(B + a.off[10], wr, total);
    B[a.off[11] + wr] = (unsigned char)is_code;
    B[a.off[12] + wr] = 0;
    ((u64*)(B + a.off[13]))[wr] = e.rw_counter + rwc;
    ((u64*)(B + a.off[14]))[wr] = rw_total - rwc;
    B[a.off[15] + wr] = e.dst_tag == CPT_MEMORY;
    B[a.off[16] + wr] = e.dst_tag == CPT_BYTECODE;
    B[a.off[17] + wr] = e.dst_tag == CPT_TXCALLDATA;
    B[a.off[18] + wr] = e.dst_tag == CPT_TXLOG;
    B[a.off[19] + wr] = e.dst_tag == CPT_RLCACC;
    a.flags[wr] = dst_word;
  }
}

#ifdef __CUDACC__
__global__ void __launch_bounds__(256) k_seg_local(SegHorner s, Fr* chunk_val) {
  const u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < s.n_chunks) seg_local(s, chunk_val, c);
}
__global__ void __launch_bounds__(128) k_seg_carry(SegHorner s, Fr* chunk_val, Fr* seg_total) {
  const u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < s.n_seg) seg_carry(s, chunk_val, seg_total, k);
}
__global__ void __launch_bounds__(256) k_assign_bytecode_rlc(const __grid_constant__ BytecodeAssign a, const Fr* chunk_val) {
  const u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < a.s.n_chunks) assign_bytecode_rlc(a, chunk_val, c);
}
__global__ void __launch_bounds__(256) k_assign_bytecode_rows(const __grid_constant__ BytecodeAssign a) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < a.n_rows; r += stride) assign_bytecode_row(a, r);
}
__global__ void __launch_bounds__(256) k_assign_state_derive(const __grid_constant__ StateAssign a) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < a.n_rows; r += stride) assign_state_row(a, r);
}
__global__ void __launch_bounds__(256) k_assign_copy_rows(const __grid_constant__ CopyAssign a, const Fr* chunk_val, const Fr* seg_total) {
  const u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < a.s.n_chunks) assign_copy_chunk(a, chunk_val, seg_total, c);
}
#endif

}  // namespace zk
