// dnastore-b200: command-line front end that keeps dnastore's flags for the Viterbi
// path and calls the GPU decoder through the C ABI (include/dnastore_b200.h).
//
// Flag names, defaults and meaning follow the reference's main
// (reference t/dnastore.cpp:41-82,115-130,150-176,217-223):
//   -l/--length N (12)            k-mer length; sets maxDupLen = N/2
//   -L/--load-machine FILE        machine JSON (plain or .gz)
//   -C/--compose-machine FILE     repeatable; first listed = outermost
//   -S/--save-machine FILE|-      write the (composed) machine JSON
//   -V/--decode-viterbi FASTA     batched Viterbi decode on the GPU
//   -d/--decode-file FASTA        exact (error-free) decode on the host, bytes to stdout (decoder.h)
//   -D/--decode-string DNA        the same for one DNA string
//   -B/--decode-bits DNA          the decoded input-symbol string (^0101...$) instead of bytes
//   -r/--raw                      print bare decoded strings, one per line
//   --error-sub-prob (.01) --error-iv-ratio (10) --error-dup-prob (.001)
//   --error-del-open (.001) --error-del-ext (.01) --error-global  -F/--error-file
//   -v/--verbose N                accepted (0 = quiet); >=2 prints decoder info to stderr
//   --device N                    CUDA ordinal (new)
//   --devices A,B,...             decode on several GPUs of this node, reads sharded over them, output in input order (new)
//   -f/--fit-error STK            Baum-Welch fit of the error model on a Stockholm database of pairwise alignments (pair-HMM
//                                 forward-backward on the GPU), fitted parameters as JSON on stdout (t/dnastore.cpp:135-140)
//   --error-counts STK            posterior expected counts of the error events as JSON on stdout (t/dnastore.cpp:142-146)
//   --strict-guides               treat the alignments as strict truth, not hints
// Machine construction (-l without -L) and the encoder (-e/-E/-b) are outside this build's scope and are reported as such.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/dnastore_b200.h"

static void die(const std::string& msg, int code = 1) {
  std::cerr << msg << std::endl;
  std::exit(code);
}

int main(int argc, char** argv) {
  dnab_error_flags ef;
  dnab_error_flags_default(&ef);
  std::string loadMachine, saveMachine, viterbiFile, errorFile, exactFile, exactString, exactBits, fitFile, countsFile;
  bool strictGuides = false;
  bool haveExactString = false, haveExactBits = false;
  std::vector<std::string> composes;
  bool raw = false;
  int verbose = 2, device = 0;
  std::vector<int> devices;

  auto longName = [](const std::string& a, std::string& name, std::string& val, bool& hasVal) {
    name = a.substr(2);
    const size_t eq = name.find('=');
    hasVal = eq != std::string::npos;
    if (hasVal) {
      val = name.substr(eq + 1);
      name = name.substr(0, eq);
    }
  };
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i], name, val;
    bool hasVal = false;
    if (a.size() > 2 && a[0] == '-' && a[1] == '-')
      longName(a, name, val, hasVal);
    else if (a.size() >= 2 && a[0] == '-') {
      static const struct { char s; const char* l; } shorts[] = {
          {'l', "length"}, {'L', "load-machine"}, {'C', "compose-machine"}, {'S', "save-machine"},
          {'V', "decode-viterbi"}, {'r', "raw"}, {'F', "error-file"}, {'v', "verbose"}, {'h', "help"},
          {'d', "decode-file"}, {'e', "encode-file"}, {'E', "encode-string"}, {'D', "decode-string"},
          {'b', "encode-bits"}, {'B', "decode-bits"}, {'f', "fit-error"}};
      for (const auto& s : shorts)
        if (s.s == a[1]) name = s.l;
      if (name.empty()) die("unrecognised option '" + a + "'");
      if (a.size() > 2) {
        val = a.substr(2);
        hasVal = true;
      }
    } else
      die("too many positional options have been specified on the command line");
    auto need = [&]() -> std::string {
      if (hasVal) return val;
      if (i + 1 >= argc) die("the required argument for option '--" + name + "' is missing");
      return argv[++i];
    };
    if (name == "help") {
      std::cout << "dnastore-b200: GPU Viterbi decoder behind dnastore's -V path; see the header of cli_main.cpp\n";
      return 1;
    } else if (name == "length") ef.length = std::atoi(need().c_str());
    else if (name == "load-machine") loadMachine = need();
    else if (name == "compose-machine") composes.push_back(need());
    else if (name == "save-machine") saveMachine = need();
    else if (name == "decode-viterbi") viterbiFile = need();
    else if (name == "decode-file") exactFile = need();
    else if (name == "decode-string") { exactString = need(); haveExactString = true; }
    else if (name == "decode-bits") { exactBits = need(); haveExactBits = true; }
    else if (name == "raw") raw = true;
    else if (name == "error-sub-prob") ef.sub_prob = std::atof(need().c_str());
    else if (name == "error-iv-ratio") ef.iv_ratio = std::atof(need().c_str());
    else if (name == "error-dup-prob") ef.dup_prob = std::atof(need().c_str());
    else if (name == "error-del-open") ef.del_open = std::atof(need().c_str());
    else if (name == "error-del-ext") ef.del_ext = std::atof(need().c_str());
    else if (name == "error-global") ef.global = 1;
    else if (name == "error-file") errorFile = need();
    else if (name == "fit-error") fitFile = need();
    else if (name == "error-counts") countsFile = need();
    else if (name == "strict-guides") strictGuides = true;
    else if (name == "verbose") verbose = std::atoi(need().c_str());
    else if (name == "device") device = std::atoi(need().c_str());
    else if (name == "devices") {
      std::string list = need(), item;
      for (size_t p = 0; p <= list.size(); ++p) {
        if (p == list.size() || list[p] == ',') {
          if (!item.empty()) devices.push_back(std::atoi(item.c_str()));
          item.clear();
        } else
          item.push_back(list[p]);
      }
      if (devices.empty()) die("--devices needs a comma-separated list of CUDA ordinals");
    }
    else if (name == "nocolor") {}
    else if (name == "log") (void)need();
    else
      die("option '--" + name + "' is outside the scope of dnastore-b200 (Viterbi decoding path only)", 2);
  }
  if (ef.length > 31) die("Maximum context is 31 bases");
  // error-model training comes first and needs no machine (t/dnastore.cpp:135-149)
  if (!fitFile.empty() || !countsFile.empty()) {
    if (!errorFile.empty()) die("--error-file together with --fit-error / --error-counts is outside the scope of dnastore-b200", 2);
    dnab_mutator_params params;
    dnab_mutator_params_from_flags(&ef, &params);
    dnab_pair_db* db = dnab_pair_db_load((fitFile.empty() ? countsFile : fitFile).c_str());
    if (!db) die(dnab_last_error());
    char* text = nullptr;
    if (!fitFile.empty()) {
      dnab_mutator_params fitted;
      int32_t iters = 0;
      if (dnab_baum_welch(device, &params, db, strictGuides ? 1 : 0, &fitted, &iters) != DNAB_OK) die(dnab_last_error(), 3);
      text = dnab_mutator_params_json(&fitted);
    } else {
      dnab_mutator_counts counts;
      double ll = 0;
      if (dnab_expected_counts(device, &params, db, strictGuides ? 1 : 0, &counts, &ll) != DNAB_OK) die(dnab_last_error(), 3);
      text = dnab_mutator_counts_json(&counts);
    }
    if (!text) die(dnab_last_error());
    std::fputs(text, stdout);
    dnab_free(text);
    dnab_pair_db_free(db);
    return 0;
  }
  if (loadMachine.empty()) die("dnastore-b200 needs --load-machine (machine construction is out of scope; "
                               "machines are reproducible only as JSON)", 2);

  dnab_machine* machine = dnab_machine_load(loadMachine.c_str());
  if (!machine) die(dnab_last_error());
  // compose arguments are applied right to left, so the first listed is outermost
  for (auto it = composes.rbegin(); it != composes.rend(); ++it) {
    dnab_machine* outer = dnab_machine_load(it->c_str());
    if (!outer) die(dnab_last_error());
    dnab_machine* composed = dnab_machine_compose(outer, machine);
    if (!composed) die(dnab_last_error());
    dnab_machine_free(outer);
    dnab_machine_free(machine);
    machine = composed;
  }

  if (!saveMachine.empty()) {
    char* text = dnab_machine_to_json(machine);
    if (!text) die(dnab_last_error());
    if (saveMachine == "-")
      std::fputs(text, stdout);
    else {
      FILE* f = std::fopen(saveMachine.c_str(), "w");
      if (!f) die("cannot write " + saveMachine);
      std::fputs(text, f);
      std::fclose(f);
    }
    dnab_free(text);
  }

  // exact decoding on the host (reference t/dnastore.cpp:185-211); same precedence as the reference's if-chain
  auto warn = [](char* lines) {
    if (!lines) return;
    std::string all = lines, line;
    dnab_free(lines);
    for (char c : all) {
      if (c == '\n') {
        std::cerr << "Warning: " << line << std::endl;
        line.clear();
      } else
        line.push_back(c);
    }
  };
  auto exactSymbols = [&](const std::string& dna) -> std::string {
    dnab_exact_decoder* xd = dnab_exact_decoder_create(machine);
    if (!xd) die(dnab_last_error());
    if (dnab_exact_decoder_feed(xd, dna.data(), dna.size()) != DNAB_OK || dnab_exact_decoder_close(xd) != DNAB_OK) {
      std::cerr << dnab_last_error() << std::endl;
      std::abort();  // the reference's Assert aborts (util.h:31)
    }
    warn(dnab_exact_decoder_warnings(xd));
    char* sym = dnab_exact_decoder_take_symbols(xd);
    std::string out = sym ? sym : "";
    dnab_free(sym);
    dnab_exact_decoder_destroy(xd);
    return out;
  };
  if (!exactFile.empty()) {
    uint8_t* bytes = nullptr;
    size_t n = 0;
    char* w = nullptr;
    const int rc = dnab_exact_decode_fasta(machine, exactFile.c_str(), &bytes, &n, &w);
    if (rc != DNAB_OK) {
      std::cerr << dnab_last_error() << std::endl;
      if (rc == DNAB_EIO) return 1;
      std::abort();
    }
    std::fwrite(bytes, 1, n, stdout);
    std::fflush(stdout);
    dnab_free(bytes);
    warn(w);
  } else if (haveExactString) {
    const std::string sym = exactSymbols(exactString);
    std::string bytes(sym.size() / 8 + 1, '\0');
    char left[16];
    char* w = nullptr;
    const int64_t n = dnab_pack_decoded_symbols(sym.data(), sym.size(), (uint8_t*)&bytes[0], bytes.size(), left, &w);
    if (n < 0) die(dnab_last_error());
    std::fwrite(bytes.data(), 1, (size_t)n, stdout);
    std::fflush(stdout);
    warn(w);
  } else if (haveExactBits) {
    std::cout << exactSymbols(exactBits) << std::endl;
  } else if (!viterbiFile.empty()) {
    dnab_compiled* compiled = errorFile.empty() ? dnab_compile(machine, &ef)
                                                : dnab_compile_with_error_file(machine, errorFile.c_str());
    if (!compiled) {
      // a null cycle is reported and the reference still exits 0 (t/dnastore.cpp:245-250)
      const std::string msg = dnab_last_error();
      std::cerr << msg << std::endl;
      return msg.find("cyclic") != std::string::npos ? 0 : 1;
    }
    if (devices.empty()) devices.push_back(device);
    dnab_multi_decoder* multi = dnab_multi_decoder_create(dnab_compiled_tables(compiled), devices.data(), (int)devices.size());
    if (!multi) die(dnab_last_error(), 3);
    dnab_decoder* dec = dnab_multi_decoder_at(multi, 0);
    if (verbose >= 3) {
      dnab_decoder_info info;
      if (dnab_decoder_get_info(dec, &info) == DNAB_OK)
        std::cerr << "decoder: " << info.n_states << " states, k=" << info.k << ", cluster of " << info.cluster_size
                  << " CTAs x " << info.threads_per_cta << " threads, " << info.states_per_cta << " states/CTA, "
                  << info.smem_bytes_per_cta << " B smem, " << info.n_clusters << " reads in flight" << std::endl;
    }
    dnab_decoded_set* out = dnab_decode_fasta_multi(multi, viterbiFile.c_str());
    if (!out) die(dnab_last_error(), 3);
    const int64_t n = dnab_decoded_count(out);
    for (int64_t i = 0; i < n; ++i) {
      if (dnab_decoded_status(out, i) == DNAB_READ_NO_DECODING) std::cerr << "No valid Viterbi decoding found" << std::endl;
      const std::string seq = dnab_decoded_seq(out, i);
      if (raw)
        std::cout << seq << "\n";
      else {
        std::cout << '>' << dnab_decoded_name(out, i) << "\n";
        for (size_t p = 0; p < seq.size(); p += 50) std::cout << seq.substr(p, 50) << "\n";
      }
    }
    dnab_decoded_free(out);
    if (verbose >= 3) {
      dnab_pipeline_stats ps;
      if (dnab_multi_decoder_last_stats(multi, &ps) == DNAB_OK)
        std::cerr << "pipeline: " << ps.reads << " reads in " << ps.chunks << " chunks on " << devices.size() << " device(s), parse "
                  << ps.parse_seconds << " s, decode " << ps.decode_busy_seconds << " s (summed), wall " << ps.wall_seconds << " s, "
                  << ps.overflow_reruns << " overflow re-decodes" << std::endl;
    }
    dnab_multi_decoder_destroy(multi);
    dnab_compiled_free(compiled);
  }
  dnab_machine_free(machine);
  return 0;
}
