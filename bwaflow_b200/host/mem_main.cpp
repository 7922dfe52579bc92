// bwaflow_mem -- the compute_flow of src/main.cpp:295-357 with every compute slot on the GPU: what `bwa-flow mem --use_gpu` runs.
//   FastqSource (KseqsRead + KseqsToBseqs) -> SeqsToChainsCUDA -> ChainsPipe -> ChainsToRegionsCUDA -> RegionsToSamCUDA -> SamsReorder -> SamsPrint
// The switch sits where --use_fpga sits in the reference (src/config.cpp:135-136, src/main.cpp:201-231): with --use_gpu the CUDA
// environment is created at start-up (index replica + device batches on every GPU) and the CUDA stage classes fill the slots
// exclusively, as the FPGA stages do under --disable_smem_cpu / --disable_sw_cpu (src/main.cpp:303-335).  This binary carries no CPU
// stages at all: without a GPU it stops with an error instead of falling back.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <string>
#include "CUDAPipeline.h"
#include "IndexFiles.h"
#include "Ingest.h"

static void usage(const char *a0)
{
  fprintf(stderr,
          "usage: %s [options] <idxbase> <in1.fq[.gz]> [in2.fq[.gz]]\n"
          "  --use_gpu[=N]      run the compute stages on N GPUs (default: all visible); this build has no CPU stages\n"
          "  -t INT             worker threads per CUDA stage (default 2 per GPU)\n"
          "  -o FILE            SAM output (default stdout);  --inorder_output  keep the input order\n"
          "  -p                 in1 holds interleaved pairs;  -C  append the FASTQ comment;  -R STR  read group line ('@RG\\tID:..')\n"
          "  -k -w -c -d -A -B -O -E -L -U -T INT   as in bwa mem;  -M/--noM  -a -Y -V -S -P  flags as in bwa mem (-M on by default)\n"
          "  --max_reads INT    reads per record the device batches are sized for (default 400000);  --chunk_size INT  bases per record (10000000)\n", a0);
}

// bwa_escape + the ID extraction of bwa_set_rg (bwa/bwa.c:403-449)
static bool set_rg(const std::string &arg, std::string &line, std::string &id)
{
  if (arg.compare(0, 3, "@RG") != 0 || arg.find('\t') != std::string::npos) return false;
  line.clear();
  for (size_t i = 0; i < arg.size(); ++i) {
    if (arg[i] == '\\' && i + 1 < arg.size()) {
      ++i;
      if (arg[i] == 't') line.push_back('\t');
      else if (arg[i] == 'n') line.push_back('\n');
      else if (arg[i] == 'r') line.push_back('\r');
      else if (arg[i] == '\\') line.push_back('\\');
      else { line.push_back('\\'); line.push_back(arg[i]); }
    } else line.push_back(arg[i]);
  }
  const size_t p = line.find("\tID:");
  if (p == std::string::npos) return false;
  size_t q = p + 4;
  while (q < line.size() && line[q] != '\t' && line[q] != '\n') ++q;
  if (q - (p + 4) + 1 > 256) return false;
  id = line.substr(p + 4, q - (p + 4));
  return true;
}

int main(int argc, char **argv)
{
  bwaflow_opt_t opt; bwaflow_opt_default(&opt);
  bwaflow_chain_opt_t copt; bwaflow_chain_opt_default(&copt);
  bwaflow_pe_opt_t po; bwaflow_pe_opt_default(&po);
  bwaflow_sam_opt_t so; bwaflow_sam_opt_default(&so);
  int n_gpus = 0, n_workers = 0, A = 0, B = 0, O = 0, E = 0, L = 0, U = 0, T = 0, zd = 0;
  bool use_gpu = false, inorder = false, interleaved = false, copy_comment = false;
  int64_t max_reads = 400000;
  int chunk = 10000000;                                       // bases per record, src/Pipeline.cpp:146
  const char *out_path = nullptr;
  std::string rg_line, rg_id, cl;
  std::vector<const char *> pos;
  for (int i = 0; i < argc; ++i) { if (i) cl += ' '; cl += argv[i]; }
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto ival = [&](int &v) { if (i + 1 >= argc) { usage(argv[0]); exit(2); } v = atoi(argv[++i]); };
    if (a == "--use_gpu") use_gpu = true;
    else if (a.compare(0, 10, "--use_gpu=") == 0) { use_gpu = true; n_gpus = atoi(a.c_str() + 10); }
    else if (a == "--inorder_output") inorder = true;
    else if (a == "--noM") so.flag &= ~0x10;
    else if (a == "--max_reads") { int v; ival(v); max_reads = v; }
    else if (a == "--chunk_size") ival(chunk);
    else if (a == "-M") so.flag |= 0x10;
    else if (a == "-a") so.flag |= 0x8;
    else if (a == "-Y") so.flag |= 0x200;
    else if (a == "-V") so.flag |= 0x100;
    else if (a == "-S") so.flag |= 0x20;
    else if (a == "-P") so.flag |= 0x4;
    else if (a == "-p") interleaved = true;
    else if (a == "-C") copy_comment = true;
    else if (a == "-t") ival(n_workers);
    else if (a == "-k") { ival(opt.min_seed_len); copt.min_seed_len = opt.min_seed_len; }
    else if (a == "-w") { ival(opt.w); copt.w = opt.w; }
    else if (a == "-c") { ival(opt.max_occ); copt.max_occ = opt.max_occ; }
    else if (a == "-d") ival(zd);
    else if (a == "-A") ival(A);
    else if (a == "-B") ival(B);
    else if (a == "-O") ival(O);
    else if (a == "-E") ival(E);
    else if (a == "-L") ival(L);
    else if (a == "-U") ival(U);
    else if (a == "-T") ival(T);
    else if (a == "-o" || a == "--output") { if (i + 1 >= argc) { usage(argv[0]); return 2; } out_path = argv[++i]; }
    else if (a == "-R") {
      if (i + 1 >= argc || !set_rg(argv[++i], rg_line, rg_id)) { fprintf(stderr, "[E::%s] bad read group line\n", argv[0]); return 2; }
    } else if (a.size() > 1 && a[0] == '-') { fprintf(stderr, "unknown option %s\n", a.c_str()); usage(argv[0]); return 2; }
    else pos.push_back(argv[i]);
  }
  if (pos.size() < 2 || pos.size() > 3) { usage(argv[0]); return 2; }
  if (!use_gpu) { fprintf(stderr, "%s: --use_gpu is required: this build carries no CPU compute stages\n", argv[0]); return 2; }
  // -A scales the other scores unless they are given (bwa/fastmap.c:290-303)
  if (A) { opt.a = A; opt.b *= A; opt.o_del *= A; opt.e_del *= A; opt.o_ins *= A; opt.e_ins *= A; opt.zdrop *= A; opt.pen_clip5 *= A; opt.pen_clip3 *= A; po.pen_unpaired *= A; so.T *= A; }
  if (B) opt.b = B;
  if (O) opt.o_del = opt.o_ins = O;
  if (E) opt.e_del = opt.e_ins = E;
  if (L) opt.pen_clip5 = opt.pen_clip3 = L;
  if (U) po.pen_unpaired = U;
  if (T) so.T = T;
  if (zd) opt.zdrop = zd;
  for (int i = 0, k = 0; i < 4; ++i) {                        // bwa_fill_scmat, bwa/bwa.c:209-220
    for (int j = 0; j < 4; ++j) opt.mat[k++] = (int8_t)(i == j ? opt.a : -opt.b);
    opt.mat[k++] = -1;
  }
  for (int j = 0; j < 5; ++j) opt.mat[20 + j] = -1;
  const bool paired = pos.size() == 3 || interleaved;
  if (paired) so.flag |= 0x2; else so.flag &= ~0x2;
  int rc = 0;
  try {
    if (bwaflow_device_count() <= 0) throw std::runtime_error("--use_gpu: no CUDA device visible (there is no CPU fallback)");
    const auto t0 = std::chrono::steady_clock::now();
    IndexFiles idx;
    idx.load(pos[0]);
    if (n_gpus <= 0 || n_gpus > bwaflow_device_count()) n_gpus = bwaflow_device_count();
    if (n_workers <= 0) n_workers = 2 * n_gpus;
    const int64_t max_bases = chunk + 4096 + 16;
    CUDAEnv cuda_env(&idx.ix, &opt, &copt, idx.is_alt.data(), n_gpus, 2 * ((n_workers + n_gpus - 1) / n_gpus), max_reads, max_bases);
    const auto t1 = std::chrono::steady_clock::now();
    FILE *out = out_path ? fopen(out_path, "w") : stdout;
    if (!out) throw std::runtime_error(std::string("cannot write ") + out_path);
    SamStageConfig cfg;
    cfg.po = po; cfg.so = so; cfg.ctg_names = idx.name_ptr.data(); cfg.ctg_is_alt = idx.is_alt_u8.data(); cfg.ctg_anno = idx.anno_ptr.data();
    cfg.rg_id = rg_id.c_str();
    so.n_threads = 0;
    {                                                           // header: bwa_format_sam_hdr, src/bwa_wrapper.cpp:25-35
      bwaflow_sam_batch_t *hb = nullptr;
      if (bwaflow_sam_batch_create(cuda_env.getPE(0).ctx, 16, 4096, &po, &so, cfg.ctg_names, cfg.ctg_is_alt, cfg.ctg_anno, cfg.rg_id, &hb))
        throw std::runtime_error(bwaflow_last_error());
      const std::string pg = "@PG\tID:bwa\tPN:bwa\tVN:0.7.17-bwaflow-cuda\tCL:" + cl;
      const int64_t n = bwaflow_sam_header(hb, rg_line.empty() ? nullptr : rg_line.c_str(), pg.c_str(), nullptr, 0);
      std::string hdr((size_t)n, '\0');
      bwaflow_sam_header(hb, rg_line.empty() ? nullptr : rg_line.c_str(), pg.c_str(), &hdr[0], n);
      fwrite(hdr.data(), 1, hdr.size(), out);
      bwaflow_sam_batch_destroy(hb);
    }
    FastqSource src(pos[1], pos.size() == 3 ? pos[2] : nullptr, copy_comment, chunk);
    SeqsToChainsCUDA seq2chain_cuda_stage(n_workers, &cuda_env);
    ChainsPipe chainpipe_stage(1, nullptr, nullptr, &copt);
    ChainsToRegionsCUDA chain2reg_cuda_stage(n_workers, &cuda_env);
    RegionsToSamCUDA reg2sam_cuda_stage(n_workers, &cuda_env, cfg);
    SamsReorder reorder_stage(inorder);
    SamsPrint print_stage(out);
    kestrelFlow::Pipeline compute_flow(7, n_workers);
    compute_flow.addStage(0, &src);
    compute_flow.addStage(1, &seq2chain_cuda_stage);
    compute_flow.addStage(2, &chainpipe_stage);
    compute_flow.addStage(3, &chain2reg_cuda_stage);
    compute_flow.addStage(4, &reg2sam_cuda_stage);
    compute_flow.addStage(5, &reorder_stage);
    compute_flow.addStage(6, &print_stage);
    compute_flow.start();
    compute_flow.wait();
    if (out != stdout) fclose(out);
    const auto t2 = std::chrono::steady_clock::now();
    const double s_idx = std::chrono::duration<double>(t1 - t0).count(), s_run = std::chrono::duration<double>(t2 - t1).count();
    fprintf(stderr, "[bwaflow_mem] %d GPU(s), %d workers per stage; index + device set-up %.2f s; %llu reads -> %llu SAM bytes in %.2f s (%.0f reads/s)\n",
            cuda_env.n_devices(), n_workers, s_idx, (unsigned long long)print_stage.n_reads, (unsigned long long)print_stage.n_bytes, s_run,
            print_stage.n_reads / (s_run > 0 ? s_run : 1));
  } catch (std::exception &e) { fprintf(stderr, "[bwaflow_mem] error: %s\n", e.what()); rc = 1; }
  return rc;
}
