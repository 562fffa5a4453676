"""Formats the `RTR <class> <job> <event> <clock>` lines a -DRES_TRACE build of csrc/solve_resident.cu prints
(class 0 = compute thread 0, 1 = MMA-issuing warp; events: 0 job start, 1 dependency met / MMA group retired,
2 group committed / compute phase done).  See tools/resident_probe.py."""
import sys
rows=[l.split() for l in open(sys.argv[1] if len(sys.argv)>1 else 'gpurun_out/rtrace.log') if l.startswith('RTR')]
names="Q0 c0 S0 PV0 c1 Q1 c2 S1 PV1 c3 Q2 c4 S2 PV2 c5 F".split()
ev={}
for _,w,j,k,c in rows:
    ev[(int(w),int(j),int(k))]=int(c)
t0=min(ev.values())
print("job   | producer: start need_done committed | compute: start mma_ready done | wait comp")
tw=tc=0
for j in range(16):
    p=[ev.get((1,j,k),0)-t0 for k in range(3)]
    c=[ev.get((0,j,k),0)-t0 for k in range(3)]
    tw+=c[1]-c[0]; tc+=c[2]-c[1]
    print("%-4s | %7d %7d %7d | %7d %7d %7d | %6d %6d"%(names[j],*p,*c,c[1]-c[0],c[2]-c[1]))
print("wait",tw,"compute",tc)
