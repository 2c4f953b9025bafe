"""
Mutation test of the native BAM reader (libncb: ncb_bam_open / ncb_bam_read_uncalled4): a small valid
Uncalled4 BAM is damaged -- bytes, bits and whole length fields of its uncompressed stream, now and
then a byte of the compressed file -- and read back.  Every outcome must be a result or an NcbError
with a message; a crash (wrong bounds check, exception across the C ABI) ends the process, which is
what tests/test_readers.py::test_damaged_bams_never_crash_the_reader looks for.

    python tests/fuzz_bam.py [seed] [files]
"""
import os
import struct
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import signal_files as SF
from nanocompore_b200 import readers as R, _lib as L
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 0)
tmp=tempfile.mkdtemp()
refs=[('tx',120),('ty',80)]
reads=[SF.make_uncalled4_read(rng,120,'RNA002',f'r{k}',ref=0) for k in range(6)]+[SF.make_uncalled4_read(rng,80,'RNA002',f's{k}',ref=1) for k in range(3)]
# build the uncompressed payload like write_bam, then mutate the PAYLOAD (so blocks still inflate) or the file bytes
text=b'@HD\tVN:1.6\tSO:coordinate\n'
out=b'BAM\1'+struct.pack('<i',len(text))+text+struct.pack('<i',len(refs))
for name,length in refs: out+=struct.pack('<i',len(name)+1)+name.encode()+b'\0'+struct.pack('<i',length)
for r in reads:
    tags=SF._aux_array('ur','i',r['ur'])+SF._aux_array('uc','s',r['uc'])+SF._aux_array('ul','s',r['ul'])
    out+=SF._record(r['ref'],0,r['name'],0,tags)
n_ok=n_err=0
for it in range(int(sys.argv[2]) if len(sys.argv)>2 else 400):
    b=bytearray(out)
    mode=rng.integers(0,3)
    for _ in range(int(rng.integers(1,4))):
        i=int(rng.integers(4,len(b)))
        if mode==0: b[i]=int(rng.integers(0,256))
        elif mode==1: b[i:i+4]=struct.pack('<i',int(rng.choice([-1,-2**31,2**31-1,0,65536,1<<20])))
        else: b[i]^=1<<int(rng.integers(0,8))
    p=os.path.join(tmp,f'f{it}.bam')
    with open(p,'wb') as f:
        for i in range(0,len(b),700): f.write(SF._bgzf_block(bytes(b[i:i+700])))
        f.write(SF._bgzf_block(b''))
    if rng.random()<0.2:   # also damage the compressed file itself
        raw=bytearray(open(p,'rb').read()); j=int(rng.integers(0,len(raw))); raw[j]=int(rng.integers(0,256)); open(p,'wb').write(raw)
    try:
        bam=R.NativeBam(p,threads=2)
        for name,length in refs:
            bam.read_uncalled4(name,length,'RNA002')
        bam.close(); n_ok+=1
    except L.NcbError:
        n_err+=1
    os.remove(p)
print('survived',n_ok,'errors',n_err)
