import itertools
best=[]
for PX in range(11,17):
  for PY in range(7,10):
    PV=PX*PY*4
    for pad in range(0,32):
      CS=PV+pad
      s=[1,PX,PX*PY]
      tot=0
      for b in range(2):
        for k in range(3):
          cnt={}
          for g in range(24):
            comp=g>>3; a=(g>>1)&3; bh=g&1
            ai=(comp+1)%3; aj=(comp+2)%3
            addr=comp*CS + a*s[ai] + (2*bh+b)*s[aj] + k*s[comp]
            cnt[addr%32]=cnt.get(addr%32,0)+1
          tot+=max(cnt.values())
      best.append((tot,3*CS,PX,PY,pad))
best.sort()
print(best[:15])
# current mapping for reference: lanes (slot, comp, ah, bh), PX=11,PY=7
