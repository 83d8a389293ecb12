import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import dnab_testutil as util, dnastore_b200 as d
c = util.compiled_for(["l4c4"], dict(length=4), True)
dec = d.Decoder(c)
print(dec.viterbi(["TGTCACGTACGTAGCA"]))
