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
b=[(37,40,'point_valid'),(72,94,'FlagScan (ballot prefix sums)'),(121,134,'window_cov (pointClassify mean + covariance)'),(141,145,'window_not_line (eigen-solve early-out)'),(148,165,'window_line (eigen-solve + 0.08 m test)'),(167,175,'cos_angle / sq_diff'),(180,270,'bitonic sort of the voxel runs'),(318,343,'ring row bulk copy + wait'),(344,364,'ordered compaction'),(365,393,'init / tag'),(396,464,'mask: events + replay'),(466,479,'curvature'),(481,515,'region bounds / region of a cell'),(523,568,'candidate list + window list'),(570,631,'pass 1 (warp 0, register-resident arg-min picks)'),(632,649,'curvature rank by counting (warps 1-15)'),(650,688,'pointClassify driver (warps 1-15) incl. barrier waits'),(689,716,'pass 2 flags + prefixes'),(717,777,'pass 3 prefixes + list bases'),(778,815,'placement'),(816,836,'outputs'),(837,878,'voxel filter: bounding box'),(879,922,'voxel filter: indices + runs'),(923,956,'voxel filter: keys + expand'),(957,994,'voxel filter: heads + centroids')]
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
