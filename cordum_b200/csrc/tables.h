// tables.h — layout of the compiled tables and encoded job columns, shared by the
// host compiler/encoder (C++) and the CUDA kernels.  Plain C structs only.
//
// Bit-parallel first-match: for every dictionary value v of every scalar
// attribute a there is a PASS-ROW: an R-bit vector whose bit r says "rule r's predicate on a
// passes for v" (vacuous predicates -> 1).  A job's candidate rules = AND of the rows its
// attribute values select; the lowest set bit is Go's first match (safety_policy.go:192-204).
// Rows are padded to whole SEGMENTS of 1024 rules = 128 B = one cache line = 8 lanes x uint4,
// which is the unit the kernel scans in rule order with early exit.
#pragma once
#include <stdint.h>

/* 16-byte units (the kernels reinterpret these as uint4 / int4) */
typedef struct __attribute__((aligned(16))) Row16 { uint32_t w[4]; } Row16;
typedef struct __attribute__((aligned(16))) Load16 { int32_t active, max_parallel; float cpu, gpu; } Load16;

#define CORDUM_SEG_RULES 1024u          /* rules per segment                                  */
#define CORDUM_SEG_U4 8u                /* uint4 per row per segment (8 lanes x 16 B = 128 B) */
#define CORDUM_GROUP 8u                 /* lanes cooperating on one job                       */

/* dictionary ids common to every attribute */
#define CORDUM_ID_EMPTY 0u              /* raw value == ""  (never matches a non-vacuous list, safety_policy.go:297) */
#define CORDUM_ID_OTHER 1u              /* non-empty but referenced by no list                                         */

/* job flag word */
#define JF_COMBO_MASK 0x7u              /* actor_type id (0 "",1 human,2 service) * 2 + secrets_present */
#define JF_MCP_USED 0x8u
#define JF_HAS_LABELS 0x10u
#define JF_APPROVED 0x20u
#define JF_REQ_NONEMPTY 0x40u           /* len(meta.requires) > 0 (strategy_least_loaded.go:228,242)   */
#define JF_REQ_UNKNOWN 0x80u            /* a non-blank requires token no pool declares                 */
#define JF_PLACE_UNSAT 0x100u           /* a placement label no worker can satisfy                     */
#define JF_TOPIC_MISSING 0x200u         /* TrimSpace(topic) == ""   (kernel.go:171)                    */
#define JF_TOPIC_UNSUPPORTED 0x400u     /* !HasPrefix(topic,"job.") (kernel.go:174)                    */
#define JF_TOPIC_RAW_EMPTY 0x800u       /* req.Topic == ""          (strategy_least_loaded.go:41)      */
#define JF_EFF_DENIED 0x1000u           /* effective config: topic matches denied_topics (kernel.go:219)*/
#define JF_EFF_NOT_ALLOWED 0x2000u      /* allowed_topics set and topic not in it (kernel.go:223)       */
#define JF_NO_REQ 0x4000u               /* the job carries no requires token any rule lists: every rule with a
                                           requires list fails containsAll (safety_policy.go:320-330)            */
#define JF_NO_LAB 0x8000u               /* no labels, or none of the label pairs any rule lists: every rule with a
                                           labels map fails labelsMatch (safety_policy.go:332-345)               */
/* row_combo index: actor_type / secrets_present combination + 6 * NO_REQ + 12 * NO_LAB (24 rows) */
#define CORDUM_COMBO_ROWS 24u
#define CORDUM_COMBO_INDEX(flags) (((flags) & JF_COMBO_MASK) + (((flags) >> 14) & 3u) * 6u)

#define CORDUM_PREF_UNKNOWN 0xFFFFFFFFu /* preferred_pool / preferred_worker_id names nothing known    */

/* Encoded job columns: column-major (one contiguous array per attribute), 96 B/job. */
typedef struct JobColumns {
  const uint32_t* tenant;       /* tenant dictionary id (fold+trim canonical)                */
  const uint32_t* tenant_pol;   /* 1 + index into policy.Tenants by EXACT string, 0 = none   */
  const uint32_t* topic;        /* topic dictionary id (keyed by the raw topic string)       */
  const uint32_t* capability;
  const uint32_t* pack;
  const uint32_t* actor;
  const uint32_t* mcp[4];       /* server, tool, resource, action                            */
  const uint32_t* pref_pool;    /* 0 none, else 1 + pool id, or CORDUM_PREF_UNKNOWN          */
  const uint32_t* pref_worker;  /* 0 none, else 1 + worker slot, or CORDUM_PREF_UNKNOWN      */
  const uint32_t* effcfg;       /* 0 none/unparsable, else effective-config id               */
  const uint64_t* risk_mask;    /* bit per referenced risk tag                               */
  const uint64_t* req_mask;     /* bit per referenced requires token (rules U pools)         */
  const uint64_t* lab_mask;     /* bit per rule label pair (k,v): labels.get(k,"") == v      */
  const uint64_t* place_lo;     /* placement need mask, bits 0..63                           */
  const uint64_t* place_hi;     /* bits 64..127                                              */
  const uint32_t* flags;        /* JF_*                                                      */
} JobColumns;
#define CORDUM_JOB_IN_BYTES 96u
#define CORDUM_JOB_OUT_BYTES 16u

/* Everything the kernels read besides the job columns.  Pointers are device pointers. */
#define CORDUM_POOL_CHUNK 512u
#define CORDUM_POOL_SORT_MAX 8192u

typedef struct DeviceTables {
  /* ---- policy */
  uint32_t n_rules, n_seg, row_u4;            /* row_u4 = n_seg * 8 uint4 per pass-row                 */
  uint32_t item_u4;                           /* uint4 (128-bit words) per scan item: a lane ANDs item_u4 * 128 rule bits
                                                 per step; tw_list holds item indices (1, 2 or 4)                  */
  const Row16* row_tenant;  uint32_t n_tenant;
  const Row16* row_topic;   uint32_t n_topic;
  const Row16* row_cap;     uint32_t n_cap;
  const Row16* row_pack;    uint32_t n_pack;
  const Row16* row_actor;   uint32_t n_actor;
  const Row16* row_combo;                     /* 6 rows: (actor_type, secrets) & alive-rule mask        */
  const Row16* row_risk;                      /* row 0: job has no referenced tag; row 1+b: tag bit b   */
  uint32_t risk_zero_row;                     /* index of an all-zero row (padding for branch-free OR loops) */
  const Row16* row_mcp[4];  uint32_t n_mcp[4];
  uint32_t mcp_ones_row[4];                   /* index of an all-ones row per MCP table (jobs without MCP labels) */
  const Row16* row_check;                     /* rules carrying a requires/labels need-mask             */
  const uint64_t* rule_req_need;              /* per rule: requires tokens it needs (subset test)       */
  const uint64_t* rule_lab_need;              /* per rule: label pairs it needs                          */
  const uint8_t* rule_dec;                    /* per rule: CORDUM_DEC_* | 0x80 if constraints non-empty */
  const uint32_t* pos2rule;                   /* bit position -> rule index: rule bits are permuted so that topic rows are sparse */
  const uint32_t* tw_off;                     /* per topic: offset of its word list in tw_list                           */
  const uint32_t* tw_cnt;                     /* per topic: number of non-zero 128-bit words of its pass-row             */
  const uint16_t* tw_list;                    /* word indices (units of Row16 within a row)                              */
  /* tenant-level MCP lists (kernel.go:190-195) and effective-config overlay (kernel.go:218-231):
     verdict per (entry, field, value id): 0 ok, 1 denied, 2 not allowed */
  const uint8_t* tenant_mcp; uint32_t mcp_stride;   /* [n_tenant_pol][4][mcp_stride]                    */
  const uint8_t* eff_mcp;                           /* [n_effcfg+1][4][mcp_stride]                      */
  const uint8_t* eff_topic; uint32_t topic_stride;  /* [n_effcfg+1][topic_stride] bit0 denied, bit1 not-allowed */
  /* ---- routing */
  const uint32_t* topic_pool_off;             /* per topic id: start in pool_list                       */
  const uint32_t* topic_pool_cnt;             /* per topic id: number of (deduplicated) pools           */
  const uint32_t* pool_list;
  const uint64_t* pool_req_mask;              /* per pool: declared requires tokens                     */
  const uint8_t* pool_req_nonempty;           /* per pool: len(pool.requires) > 0                       */
  uint64_t req_blank_mask;                    /* bit of the token that trims to "" (ignored by pools)   */
  uint32_t n_pools;
  /* ---- workers (pool-sorted order = "pos") */
  uint32_t n_pos;                             /* routable workers                                       */
  const uint32_t* pool_off;                   /* [n_pools+1] pos ranges                                 */
  const uint32_t* pos_pool;                   /* pool id of pos                                         */
  const uint32_t* pos_slot;                   /* caller slot of pos                                     */
  const uint32_t* pos_rank;                   /* rank of pos in ascending worker_id byte order          */
  const uint32_t* slot_pos;                   /* [n_slots] 1 + pos, 0 = not routable                    */
  const uint32_t* rank_slot;                  /* [n_slots] slot of rank                                 */
  const uint32_t* rank_pos;                   /* [n_slots] pos of rank (routable workers only)          */
  const uint64_t* pos_label_lo;               /* placement label mask                                   */
  const uint64_t* pos_label_hi;
  const Load16* loads;                          /* [n_slots] {active, max_parallel, cpu(f32), gpu(f32)}   */
  uint64_t* pos_key;                          /* derived: (orderable score << 32) | rank; score field all-ones = overloaded */
  uint64_t* skey;                             /* derived: per pool, keys in ascending order (load-sorted view of the pool)  */
  uint64_t* slab_lo;                          /* derived: label masks permuted into the same order                          */
  uint64_t* slab_hi;
  uint8_t* pool_sorted;                       /* derived: 1 if the pool's sorted view is valid (pool fits the sort buffer)  */
  uint32_t* pool_nok;                         /* derived: workers of the pool that are NOT overloaded (= prefix of the sorted view) */
  uint32_t* lbm;                              /* derived: label bitmaps over the sorted view: lbm[lbm_off[p] + bit*words(p) + w],
                                                 bit i of word w = "sorted worker 32w+i carries label bit"               */
  const uint32_t* lbm_off;                    /* [n_pools] word offset of the pool's bitmaps                               */
  uint32_t place_bits;                        /* label bits in use (<= 128)                                                */
  /* worker-table refresh: a pool is cut into chunks of CORDUM_POOL_CHUNK workers, one CTA sorts one chunk, a second
     kernel merges the chunks of pools that have more than one (pools above CORDUM_POOL_SORT_MAX stay unsorted)       */
  const uint32_t* chunk_pool;                 /* [n_chunks] pool of the chunk                                              */
  const uint32_t* pool_chunk0;                /* [n_pools+1] first chunk of the pool                                       */
  const uint32_t* merge_list;                 /* [n_merge] chunks worker_merge_kernel runs (all chunks of a multi-chunk
                                                 sortable pool; chunk 0 only of an unsortable pool)                       */
  uint32_t n_chunks, n_merge;
  uint32_t merge_smem;                        /* bytes of shared memory worker_merge_kernel needs (largest merged pool)    */
  uint64_t* ckey;                             /* derived scratch: per chunk, its keys in ascending order                   */
  uint64_t* pool_best;                        /* derived: min key per pool (~0 = none)                  */
  uint32_t* pool_mincnt;                      /* derived: workers in the pool sharing the min score     */
} DeviceTables;
