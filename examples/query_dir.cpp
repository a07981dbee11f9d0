// query_dir — a Go-less host for the hot path, end to end in native code:
//   sybil table directory --(libsybilgob: gob decode)--> sg_block_desc --(libsybilgpu: stage + scan)--> result
//
//   g++ -O2 -std=c++17 -I include examples/query_dir.cpp -L sybil_b200/csrc -lsybilgob -lsybilgpu -lz
//       -Wl,-rpath,$PWD/sybil_b200/csrc -o query_dir
//   ./query_dir <dbdir> <table> <group-col> <agg-col> [hist]
//
// What `sybil query -table T -group G -int A [-op hist]` does through Table.LoadAndQueryRecords
// (src/lib/table_query.go:18), minus the CLI: open the table, stage every block (only the two columns the
// query names: the LoadSpec), run, print "group key <tab> count <tab> mean" per group, Count descending.
// Needs a CUDA device: the library has no CPU path and says so.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "sybilgob.h"
#include "sybilgpu.h"

int main(int argc, char** argv) {
  if (argc < 5) {
    fprintf(stderr, "usage: %s <dbdir> <table> <group-col> <agg-col> [hist]\n", argv[0]);
    return 2;
  }
  const bool hist = argc > 5 && !strcmp(argv[5], "hist");
  char err[512];
  sgob_table* tinfo = sgob_table_open(argv[1], argv[2], err, sizeof err);
  if (!tinfo) {
    fprintf(stderr, "%s\n", err);
    return 1;
  }
  const int32_t ncols = sgob_table_num_cols(tinfo);
  std::vector<const char*> names((size_t)ncols);
  std::vector<int32_t> types((size_t)ncols);
  std::vector<uint8_t> load((size_t)ncols, 0);
  int32_t gslot = -1, aslot = -1;
  for (int32_t s = 0; s < ncols; s++) {
    names[(size_t)s] = sgob_table_col_name(tinfo, s);
    types[(size_t)s] = sgob_table_col_type(tinfo, s);
    if (!strcmp(names[(size_t)s], argv[3])) gslot = s;
    if (!strcmp(names[(size_t)s], argv[4])) aslot = s;
  }
  if (gslot < 0 || aslot < 0 || types[(size_t)aslot] != SG_COL_INT) {
    fprintf(stderr, "unknown group column or non-int aggregation column\n");
    return 1;
  }
  load[(size_t)gslot] = load[(size_t)aslot] = 1;

  int status = 0;
  sg_ctx* ctx = sg_create(0, &status);
  if (status != SG_OK) {
    fprintf(stderr, "sg_create: %s\n", sg_last_error(ctx));
    return 1;
  }
  sg_table* table = sg_table_create(ctx, ncols, types.data());
  if (!table) {
    fprintf(stderr, "sg_table_create: %s\n", sg_last_error(ctx));
    return 1;
  }
  int64_t staged = 0;
  for (int64_t i = 0; i < sgob_table_num_blocks(tinfo); i++) {
    sgob_block* b = sgob_read_block_dir(sgob_table_block_dir(tinfo, i), names.data(), types.data(), ncols, load.data(), i, err, sizeof err);
    if (!b) {
      fprintf(stderr, "skipping block: %s\n", err);  // LoadBlockFromDir returning nil (table_query.go:134-139)
      continue;
    }
    if (sg_table_add_block(table, sgob_block_desc(b)) == SG_OK) staged++;
    else fprintf(stderr, "skipping block: %s\n", sg_last_error(ctx));
    sgob_block_free(b);
  }

  sg_group_desc group;
  memset(&group, 0, sizeof group);
  group.col_slot = gslot;
  group.col_type = types[(size_t)gslot];
  sg_agg_desc agg;
  memset(&agg, 0, sizeof agg);
  agg.col_slot = aslot;
  int64_t lo = 0, hi = 0;
  sgob_table_int_info(tinfo, aslot, &lo, &hi);  // histogram extents come from the TABLE's IntInfo (hist.go:27-38)
  agg.info_min = lo;
  agg.info_max = hi;
  sg_query_desc q;
  memset(&q, 0, sizeof q);
  q.abi_version = SG_ABI_VERSION;
  q.ngroups = 1;
  q.groups = &group;
  q.naggs = 1;
  q.aggs = &agg;
  q.op_mode = hist ? SG_MODE_HIST : SG_MODE_AVG;
  q.hist_kind = SG_HIST_BASIC;
  q.time_col_slot = -1;
  q.weight_col_slot = -1;
  sg_query* query = sg_query_begin(ctx, table, &q);
  sg_result* res = nullptr;
  if (!query || sg_query_run(query) != SG_OK || sg_query_finish(query, &res) != SG_OK) {
    fprintf(stderr, "query: %s\n", sg_last_error(ctx));
    return 1;
  }
  printf("# %lld blocks staged, %lld rows matched, %lld groups\n", (long long)staged, (long long)sg_result_matched_count(res),
         (long long)sg_result_num_groups(res));
  for (int64_t g = 0; g < sg_result_num_groups(res); g++) {
    const char* key;
    int64_t len, count = 0, samples = 0;
    uint64_t words[SG_MAX_GROUPS];
    sg_result_group_key(res, g, &key, &len);
    sg_result_group(res, g, words, &count, &samples);
    sg_hist_view h;
    const bool has = sg_result_hist(res, g, 0, &h) == 1;
    printf("%.*s%lld\t%.6f\n", (int)len, key, (long long)count, has ? h.avg : 0.0);
  }
  sg_result_free(res);
  sg_query_free(query);
  sg_table_free(table);
  sg_destroy(ctx);
  sgob_table_free(tinfo);
  return 0;
}
