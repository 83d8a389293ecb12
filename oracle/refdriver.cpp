// TEST INFRASTRUCTURE ONLY.  A small driver LINKED AGAINST THE UNMODIFIED
// REFERENCE OBJECTS (compiled in place from /root/reference/src by
// oracle/Makefile into oracle/_ref/).  It exposes what the reference CLI never
// prints: the Viterbi log-likelihood, the traceback path and (optionally) every
// DP cell, all obtained through the reference's own public API
//   ViterbiMatrix(...)            viterbi.h:94, viterbi.cpp:62-176
//   ViterbiMatrix::traceback()    viterbi.cpp:195-304
//   loglike()/sCell/dCell/tCell   viterbi.h:98-102  (const overloads)
//   Machine::fromFile / compose   trans.cpp:477-482, :505-602
//   Encoder<FastaWriter>          encoder.h:239-242
// The traceback PATH is recovered from the reference's own level-9 log lines
// ("Traceback at (name,pos,mutState)", viterbi.cpp:214) by capturing std::clog.
// Nothing here is product code; nothing in the product links or runs it.
//
// usage:
//   refdriver viterbi [opts] --machine M.json [--compose C.json ...] --fasta reads.fa
//       opts: -l N, --sub p, --iv r, --dup p, --delopen p, --delext p, --global,
//             --path (emit path), --cells FILE (dump all cells of read 0, raw fp64)
//       one line per read:  name \t loglike(%.17g) \t loglike(hexfloat) \t decoded \t path
//       path = space-separated "state:pos:mut" triples in traceback order (end first)
//   refdriver encode --machine M.json [--compose ...]   (payload symbol strings on stdin,
//       one per line; one encoded DNA string per line on stdout)
//   refdriver compose --machine M.json --compose ... --save out.json
//   refdriver fb [-l N --sub p --iv r --dup p --delopen p --delext p] [--strict] --stk FILE
//       pair-HMM forward/backward (FwdBackMatrix, fwdback.cpp:118-188) per alignment:
//       idx \t fwd loglike (hexfloat) \t back loglike (hexfloat) \t counts (%.17g: nDelOpen nTanDup
//       nNoGap nDelExtend nDelEnd, nLen[maxDupLen], nSub[16] row-major)
//   refdriver fit  [...same flags...] [--strict] --stk FILE
//       baumWelchParams with the Laplace prior the CLI uses (t/dnastore.cpp:135-140); prints the
//       fitted parameters with %.17g
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "vguard.h"
#include "logger.h"
#include "kmer.h"
#include "trans.h"
#include "encoder.h"
#include "fastseq.h"
#include "mutator.h"
#include "viterbi.h"
#include "fwdback.h"

using namespace std;

static Machine loadMachine(const string& base, const vector<string>& comps) {
  Machine machine = Machine::fromFile(base.c_str());
  // t/dnastore.cpp:159-165: compose arguments applied right-to-left, first listed = outermost
  for (auto it = comps.rbegin(); it != comps.rend(); ++it)
    machine = Machine::compose(Machine::fromFile(it->c_str()), machine);
  return machine;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    cerr << "usage: refdriver viterbi|encode|compose ...\n";
    return 2;
  }
  const string mode = argv[1];
  string machineFile, fasta, cellsFile, saveFile, stkFile;
  bool strict = false;
  vector<string> comps;
  int len = 12;
  double sub = .01, iv = 10, dup = .001, delOpen = .001, delExt = .01;
  bool global = false, wantPath = false;
  for (int i = 2; i < argc; ++i) {
    const string a = argv[i];
    auto next = [&]() -> string {
      if (i + 1 >= argc) {
        cerr << "missing value for " << a << endl;
        exit(2);
      }
      return string(argv[++i]);
    };
    if (a == "--machine") machineFile = next();
    else if (a == "--compose") comps.push_back(next());
    else if (a == "--fasta") fasta = next();
    else if (a == "--cells") cellsFile = next();
    else if (a == "--save") saveFile = next();
    else if (a == "-l") len = atoi(next().c_str());
    else if (a == "--sub") sub = atof(next().c_str());
    else if (a == "--iv") iv = atof(next().c_str());
    else if (a == "--dup") dup = atof(next().c_str());
    else if (a == "--delopen") delOpen = atof(next().c_str());
    else if (a == "--delext") delExt = atof(next().c_str());
    else if (a == "--global") global = true;
    else if (a == "--strict") strict = true;
    else if (a == "--stk") stkFile = next();
    else if (a == "--path") wantPath = true;
    else {
      cerr << "unknown argument " << a << endl;
      return 2;
    }
  }
  logger.setVerbose(0);
  logger.colorOff();

  if (mode == "fb" || mode == "fit") {
    MutatorParams mut;  // exactly as t/dnastore.cpp:119-129
    mut.initMaxDupLen(len / 2);
    mut.pTanDup = dup;
    mut.pDelOpen = delOpen;
    mut.pDelExtend = delExt;
    mut.pTransition = sub * iv / (1 + iv);
    mut.pTransversion = sub / (1 + iv);
    mut.local = !global;
    const list<Stockholm> db = readStockholmDatabase(stkFile.c_str());
    if (mode == "fit") {
      MutatorCounts prior(mut);
      prior.initLaplace();
      const MutatorParams fit = baumWelchParams(mut, prior, db, strict);
      printf("%.17g %.17g %.17g %.17g %.17g", fit.pDelOpen, fit.pDelExtend, fit.pTanDup, fit.pTransition, fit.pTransversion);
      for (double p : fit.pLen) printf(" %.17g", p);
      printf("\n");
      return 0;
    }
    size_t idx = 0;
    for (const auto& stock : db) {
      const FwdBackMatrix fb(mut, stock, strict);
      const MutatorCounts c = fb.counts();
      printf("%zu\t%a\t%a\t%.17g %.17g %.17g %.17g %.17g", idx++, fb.fwd.loglike, fb.back.loglike, c.nDelOpen, c.nTanDup,
             c.nNoGap, c.nDelExtend, c.nDelEnd);
      for (double v : c.nLen) printf(" %.17g", v);
      for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) printf(" %.17g", c.nSub[i][j]);
      printf("\n");
    }
    return 0;
  }

  const Machine machine = loadMachine(machineFile, comps);

  if (mode == "compose") {
    ofstream out(saveFile);
    machine.writeJSON(out);
    return 0;
  }

  if (mode == "encode") {
    string line;
    while (getline(cin, line)) {
      ostringstream enc;
      {
        FastaWriter writer(enc, NULL);
        Encoder<FastaWriter> encoder(machine, writer);
        encoder.encodeSymbolString(line);
        encoder.close();
      }
      string s = enc.str();
      string flat;
      for (char c : s)
        if (c != '\n') flat.push_back(c);
      cout << flat << "\n";
    }
    return 0;
  }

  if (mode != "viterbi") {
    cerr << "unknown mode " << mode << endl;
    return 2;
  }

  // error model exactly as t/dnastore.cpp:119-129
  MutatorParams mut;
  mut.initMaxDupLen(len / 2);
  mut.pTanDup = dup;
  mut.pDelOpen = delOpen;
  mut.pDelExtend = delExt;
  mut.pTransition = sub * iv / (1 + iv);
  mut.pTransversion = sub / (1 + iv);
  mut.local = !global;

  // viterbi.cpp:306-311
  const vguard<FastSeq> reads = readFastSeqs(fasta.c_str());
  const string inAlph = machine.inputAlphabet(MachineRelaxedInputFlag | MachineControlInputFlag | MachineSEOFInputFlag);
  const InputModel inmod(inAlph, 1., pow(4., -(double)(4 * mut.maxDupLen())));

  map<string, State> nameToState;
  if (wantPath)
    for (State s = 0; s < machine.nStates(); ++s) {
      if (nameToState.count(machine.state[s].name)) {
        cerr << "duplicate state name " << machine.state[s].name << "; cannot recover path" << endl;
        return 3;
      }
      nameToState[machine.state[s].name] = s;
    }

  size_t readIdx = 0;
  for (const auto& read : reads) {
    const ViterbiMatrix vit(machine, inmod, mut, read);
    const ViterbiMatrix& cvit = vit;
    const double ll = cvit.loglike();

    string decoded, pathStr;
    if (wantPath) {
      ostringstream cap;
      streambuf* old = clog.rdbuf(cap.rdbuf());
      logger.setVerbose(9);
      decoded = cvit.traceback();
      logger.setVerbose(0);
      clog.rdbuf(old);
      istringstream lines(cap.str());
      string ln;
      const string key = "Traceback at (";
      while (getline(lines, ln)) {
        const size_t p = ln.find(key);
        if (p == string::npos) continue;
        string body = ln.substr(p + key.size());
        const size_t close = body.rfind(')');
        body = body.substr(0, close);
        const size_t c2 = body.rfind(',');
        const size_t c1 = body.rfind(',', c2 - 1);
        const string nm = body.substr(0, c1), posS = body.substr(c1 + 1, c2 - c1 - 1), ms = body.substr(c2 + 1);
        int mutIdx = ms == "S" ? 0 : (ms == "D" ? 1 : 1 + atoi(ms.c_str() + 1));
        if (!pathStr.empty()) pathStr.push_back(' ');
        pathStr += to_string(nameToState.at(nm)) + ":" + posS + ":" + to_string(mutIdx);
      }
    } else
      decoded = cvit.traceback();

    char buf[128];
    snprintf(buf, sizeof buf, "%.17g\t%a", ll, ll);
    cout << read.name << "\t" << buf << "\t" << decoded << "\t" << pathStr << "\n";

    if (readIdx == 0 && !cellsFile.empty()) {
      // layout [pos][state][S, D, T1..Tk] with k = min(maxLeftContext, len/2) (viterbi.h:65-67)
      const size_t k = min(machine.maxLeftContext(), mut.maxDupLen());
      ofstream out(cellsFile, ios::binary);
      for (Pos pos = 0; pos <= (Pos)read.length(); ++pos)
        for (State s = 0; s < machine.nStates(); ++s) {
          double v = cvit.sCell(s, pos);
          out.write((const char*)&v, 8);
          v = cvit.dCell(s, pos);
          out.write((const char*)&v, 8);
          for (size_t i = 0; i < k; ++i) {
            v = cvit.tCell(s, pos, i);
            out.write((const char*)&v, 8);
          }
        }
    }
    ++readIdx;
  }
  return 0;
}
