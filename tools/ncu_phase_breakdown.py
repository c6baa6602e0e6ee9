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
b=[(36,39,'point_valid'),(71,93,'FlagScan (ballot prefix sums)'),(120,133,'window_cov (pointClassify mean + covariance)'),(140,144,'window_not_line (eigen-solve early-out)'),(147,164,'window_line (eigen-solve + 0.08 m test)'),(166,174,'cos_angle / sq_diff'),(179,269,'bitonic sort of the voxel runs'),(317,342,'ring row bulk copy + wait'),(343,363,'ordered compaction'),(364,392,'init / tag'),(395,463,'mask: events + replay'),(465,478,'curvature'),(480,512,'region bounds / region of a cell'),(520,566,'candidate list + window list'),(569,599,'pass 1 (warp 0, chained arg-min picks)'),(600,638,'pointClassify driver (warps 1-15) incl. barrier waits'),(642,655,'curvature rank by counting'),(656,683,'pass 2 flags + prefixes'),(684,744,'pass 3 prefixes + list bases'),(745,781,'placement'),(783,802,'outputs'),(804,845,'voxel filter: bounding box'),(846,889,'voxel filter: indices + runs'),(890,923,'voxel filter: keys + expand'),(924,960,'voxel filter: heads + centroids')]
acc=collections.Counter(); accs=collections.Counter()
for (f,l),v in per.items():
    if f=='cm_scanreg.cu':
        name=None
        for lo,hi,n in b:
            if lo<=l<=hi: name=n
        if name is None: name='cm_scanreg.cu other (l%d)'%(l//20*20)
    else: name=f + (' (eig3_sym / givens: the 3x3 eigen-solves of pointClassify)' if f == 'cm_math.h' else '')
    acc[name]+=v; accs[name]+=smp[(f,l)]
for n,v in acc.most_common(40): print("%5.1f%% inst %5.1f%% smp  %s"%(100*v/tot,100*accs[n]/tots,n))
