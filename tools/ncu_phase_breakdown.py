#!/usr/bin/env python
"""Instruction and stall-sample share of sr_ring_kernel per PHASE (source line ranges of cm_scanreg.cu; inlined helpers are
listed by file) from an `ncu --set full --import-source on` report.
usage: tools/ncu_phase_breakdown.py report.ncu-rep"""
import subprocess, csv, io, collections, sys
rep = sys.argv[1]
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass","--kernel-name","regex:sr_ring"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
fpath=None; hdr=None; per=collections.Counter(); smp=collections.Counter()
for r in rows:
    if len(r)==2 and r[0]=="File Path": fpath=r[1]; continue
    if len(r)==2 and r[0]=="Function Name": hdr=None; continue
    if hdr is None: hdr=r; continue
    try:
        ln=r[0]
        if not ln.strip(): continue
        ins=int(r[hdr.index("Instructions Executed")] or 0); s=int(r[hdr.index("# Samples")] or 0)
    except Exception: continue
    per[(fpath.split('/')[-1], int(ln))]+=ins; smp[(fpath.split('/')[-1], int(ln))]+=s
tot=sum(per.values()); tots=sum(smp.values())
b=[(20,59,'point_valid / small helpers'),(60,99,'block_scan_excl (CTA prefix sums)'),(100,139,'classify_window (pointClassify: mean, covariance, line test)'),(140,172,'cos_angle / sq_diff'),(173,221,'load+compact'),(222,249,'init/tag'),(250,320,'mask replay'),(321,335,'curvature'),(336,392,'regions/nf lists'),(393,425,'pass1 flat pick'),(426,462,'pointClassify driver (shared windows)'),(463,476,'curvature rank sort (per region)'),(477,547,'prefix/list bases'),(548,605,'pass2/3 lists + outputs'),(606,666,'voxel bbox/idx'),(667,687,'runs'),(688,698,'bitonic sort of the voxel runs'),(699,725,'centroids')]
acc=collections.Counter(); accs=collections.Counter()
for (f,l),v in per.items():
    if f=='cm_scanreg.cu':
        name=None
        for lo,hi,n in b:
            if lo<=l<=hi: name=n
        if name is None: name='scanreg helpers <173 (l%d)'%(l//20*20)
    else: name=f + (' (eig3_sym / givens: the 3x3 eigen-solves of pointClassify)' if f == 'cm_math.h' else '')
    acc[name]+=v; accs[name]+=smp[(f,l)]
for n,v in acc.most_common(40): print("%5.1f%% inst %5.1f%% smp  %s"%(100*v/tot,100*accs[n]/tots,n))
