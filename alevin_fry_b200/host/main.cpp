// main.cpp — `alevin-fry quant` command-line clone (reference src/main.rs:294-348 for the
// flags, 633-821 for validation and dispatch). The `quant` and `infer` sub-commands exist here:
// generate-permit-list and collate stay upstream (out of scope).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <unistd.h>

#include "../../include/afq_host.h"

static const char* VERSION = "0.18.0-afq-b200";

static void usage() {
  fprintf(stderr,
          "Quantify expression from a collated RAD file (B200-native)\n\n"
          "Usage: alevin-fry quant [OPTIONS] --input-dir <INPUTDIR> --tg-map <TGMAP> --output-dir <OUTPUTDIR> --resolution <RESOLUTION>\n\n"
          "Options:\n"
          "  -i, --input-dir <INPUTDIR>      input directory containing collated RAD file\n"
          "  -m, --tg-map <TGMAP>            transcript to gene map\n"
          "  -o, --output-dir <OUTPUTDIR>    output directory where quantification results will be written\n"
          "  -t, --threads <THREADS>         number of threads to use for processing (minimum: 2; lower values use 2)\n"
          "  -d, --dump-eqclasses            flag for dumping equivalence classes\n"
          "  -b, --num-bootstraps <N>        number of bootstraps to use [default: 0]\n"
          "      --init-uniform              flag for uniform sampling\n"
          "      --summary-stat              flag for storing only summary statistics\n"
          "      --use-mtx                   write the output matrix in matrix market format (default)\n"
          "      --quant-subset <SFILE>      file containing list of barcodes to quantify\n"
          "  -r, --resolution <RESOLUTION>   trivial, cr-like, cr-like-em, parsimony, parsimony-em, parsimony-gene, parsimony-gene-em\n"
          "      --small-thresh <N>          cells with fewer records take the cr-like fast path [default: 100]\n"
          "      --device <N>                CUDA device ordinal [default: 0]\n"
          "      --devices <LIST>            comma list of CUDA ordinals (or `all`): one reader feeds every listed GPU, one matrix is written\n");
}

// `alevin-fry infer` (src/main.rs:350-365, 823-848)
static int infer_main(int argc, char** argv) {
  std::string count_mat, eq_labels, output, subset;
  unsigned threads = std::max(2u, std::thread::hardware_concurrency()), device = 0;
  bool usa = false;
  auto need = [&](int& i) -> const char* {
    if (i + 1 >= argc) { fprintf(stderr, "error: a value is required for '%s'\n", argv[i]); exit(2); }
    return argv[++i];
  };
  for (int i = 2; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "-c" || a == "--count-mat") count_mat = need(i);
    else if (a == "-e" || a == "--eq-labels") eq_labels = need(i);
    else if (a == "-o" || a == "--output-dir") output = need(i);
    else if (a == "-t" || a == "--threads") threads = (unsigned)atoi(need(i));
    else if (a == "--usa") usa = true;
    else if (a == "--quant-subset") subset = need(i);
    else if (a == "--use-mtx") {}
    else if (a == "--use-eds") { fprintf(stderr, "Error: --use-eds is no longer supported. EDS output has been removed as of v0.12.\n"); return 1; }
    else if (a == "--device") device = (unsigned)atoi(need(i));
    else if (a == "-h" || a == "--help") {
      fprintf(stderr, "Perform inference on equivalence class count data\n\nUsage: alevin-fry infer [OPTIONS] --count-mat <EQCMAT> --eq-labels <EQLABELS> --output-dir <OUTPUTDIR>\n\n"
                      "Options:\n  -c, --count-mat <EQCMAT>      matrix of cells by equivalence class counts\n  -e, --eq-labels <EQLABELS>    file containing the gene labels of the equivalence classes\n"
                      "  -o, --output-dir <OUTPUTDIR>  output directory where quantification results will be written\n  -t, --threads <THREADS>       number of threads to use for processing\n"
                      "      --usa                     flag specifying that input equivalence classes were computed in USA mode\n      --quant-subset <SFILE>    file containing list of barcodes to quantify\n"
                      "      --device <N>              CUDA device ordinal [default: 0]\n");
      return 0;
    } else { fprintf(stderr, "error: unexpected argument '%s' found\n", a.c_str()); return 2; }
  }
  if (count_mat.empty() || eq_labels.empty() || output.empty()) {
    fprintf(stderr, "error: the following required arguments were not provided: --count-mat --eq-labels --output-dir\n");
    return 2;
  }
  afqh_infer_opts o{};
  o.count_mat = count_mat.c_str(); o.eq_labels = eq_labels.c_str(); o.output_dir = output.c_str();
  o.usa_mode = usa; o.filter_list = subset.empty() ? nullptr : subset.c_str(); o.num_threads = threads < 2 ? 2 : threads; o.device = (int)device;
  char err[1024] = {0};
  if (afqh_infer(&o, err, sizeof err) != 0) { fprintf(stderr, "Error: %s\n", err); return 1; }
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 2 && strcmp(argv[1], "infer") == 0) return infer_main(argc, argv);
  if (argc < 2 || strcmp(argv[1], "quant") != 0) {
    if (argc >= 2 && (strcmp(argv[1], "--version") == 0 || strcmp(argv[1], "-V") == 0)) { printf("alevin-fry %s\n", VERSION); return 0; }
    fprintf(stderr, "only the `quant` and `infer` sub-commands are implemented in this build\n");
    usage();
    return 2;
  }
  std::string cmdline;
  for (int i = 0; i < argc; ++i) { if (i) cmdline += " "; cmdline += argv[i]; }
  std::string input, tgmap, output, res, sa = "winner-take-all", subset, devices;
  unsigned threads = std::max(2u, std::thread::hardware_concurrency());
  unsigned nboot = 0, device = 0;
  bool dump_eq = false, init_uniform = false, summary_stat = false;
  long long umi_edit = -1, large_thresh = -1;
  unsigned long long small_thresh = 100;
  auto need = [&](int& i) -> const char* {
    if (i + 1 >= argc) { fprintf(stderr, "error: a value is required for '%s'\n", argv[i]); exit(2); }
    return argv[++i];
  };
  for (int i = 2; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "-i" || a == "--input-dir") input = need(i);
    else if (a == "-m" || a == "--tg-map") tgmap = need(i);
    else if (a == "-o" || a == "--output-dir") output = need(i);
    else if (a == "-t" || a == "--threads") threads = (unsigned)atoi(need(i));
    else if (a == "-d" || a == "--dump-eqclasses") dump_eq = true;
    else if (a == "-b" || a == "--num-bootstraps") nboot = (unsigned)atoi(need(i));
    else if (a == "--init-uniform") init_uniform = true;
    else if (a == "--summary-stat") summary_stat = true;
    else if (a == "--use-mtx") {}
    else if (a == "--use-eds") { fprintf(stderr, "Error: --use-eds is no longer supported. EDS output has been removed as of v0.12.\n"); return 1; }
    else if (a == "--quant-subset") subset = need(i);
    else if (a == "-r" || a == "--resolution") res = need(i);
    else if (a == "--sa-model") sa = need(i);
    else if (a == "--umi-edit-dist") umi_edit = atoll(need(i));
    else if (a == "--large-graph-thresh") large_thresh = atoll(need(i));
    else if (a == "--small-thresh") small_thresh = strtoull(need(i), nullptr, 10);
    else if (a == "--multi-sample-output") need(i);
    else if (a == "--device") device = (unsigned)atoi(need(i));
    else if (a == "--devices") devices = need(i);
    else if (a == "-h" || a == "--help") { usage(); return 0; }
    else { fprintf(stderr, "error: unexpected argument '%s' found\n", a.c_str()); usage(); return 2; }
  }
  if (input.empty() || tgmap.empty() || output.empty() || res.empty()) {
    fprintf(stderr, "error: the following required arguments were not provided: --input-dir --tg-map --output-dir --resolution\n");
    usage();
    return 2;
  }
  std::string rl = res;
  for (auto& c : rl) c = (char)tolower((unsigned char)c);
  const bool pug = rl.rfind("parsimony", 0) == 0;
  if (threads < 2) threads = 2;  // src/utils.rs:33-47
  // default_value_ifs (src/main.rs:320-341)
  if (umi_edit < 0) umi_edit = pug ? 1 : 0;
  if (large_thresh < 0) large_thresh = pug ? 1000 : 0;
  int pug_exact = 0;
  if (umi_edit == 0) { pug_exact = pug ? 1 : 0; }
  else if (umi_edit == 1) {
    if (!pug) { fprintf(stderr, "\n\nResolution strategy %s doesn't currently support 1-edit UMI resolution\nError: Invalid command line option\n", res.c_str()); return 1; }
  } else {
    fprintf(stderr, "\n\nResolution strategy %s doesn't currently support %lld-edit UMI resolution\nError: Invalid command line option\n", res.c_str(), umi_edit);
    return 1;
  }
  if (dump_eq && rl == "trivial") { fprintf(stderr, "\n\nGene equivalence classes are not meaningful in case of Trivial resolution.\n"); return 1; }
  if (nboot > 0 && !(rl == "cr-like-em" || rl == "parsimony-em" || rl == "parsimony-gene-em")) {
    fprintf(stderr, "\n\nThe num_bootstraps argument was set to %u, but bootstrapping can only be used with the cr-like-em, parsimony-em, or parsimony-gene-em resolution strategies\n", nboot);
    return 1;
  }
  afqh_quant_opts o{};
  o.input_dir = input.c_str(); o.tg_map = tgmap.c_str(); o.output_dir = output.c_str();
  o.num_threads = threads; o.num_bootstraps = nboot;
  o.init_uniform = init_uniform; o.summary_stat = summary_stat; o.dump_eq = dump_eq;
  o.resolution = res.c_str(); o.pug_exact_umi = pug_exact; o.sa_model = sa.c_str();
  o.small_thresh = small_thresh; o.large_graph_thresh = (uint64_t)large_thresh;
  o.filter_list = subset.empty() ? nullptr : subset.c_str();
  o.cmdline = cmdline.c_str(); o.version = VERSION; o.device = (int)device;
  o.devices = devices.empty() ? nullptr : devices.c_str();
  char err[1024] = {0};
  o.process_exits = 1;
  if (afqh_quantify(&o, err, sizeof err) != 0) { fprintf(stderr, "Error: %s\n", err); return 1; }
  // every output file is closed: leave without running the static destructors (the CUDA runtime's process teardown costs
  // 0.2-0.3 s that a command-line tool has no use for)
  fflush(stdout); fflush(stderr);
  _exit(0);
}
