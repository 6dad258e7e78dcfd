import collections, itertools
G=4
def conflicts(addrs):
    banks=collections.defaultdict(set)
    for a in addrs: banks[a%32].add(a)
    return max(len(v) for v in banks.values())
def conv3(H, S, SP):
    TX=H//4; PT=(H//2)*TX
    tot=0
    for warp in range(8):
        for r in range(4):
            for x in range(6):
                ad=[]
                for lane in range(32):
                    q=warp*32+lane; pt=q%PT; s=(q//PT)%G
                    ty=(pt//TX)*2; tx=(pt%TX)*4
                    ad.append(s*S+(ty+r)*SP+tx+x)
                tot+=conflicts(ad)
    return tot/(8*24)
def convt(HIN, S, SP):
    TX=HIN//4; PT=(HIN//2)*TX
    tot=0
    for warp in range(8):
        for r in range(3):
            for x in range(5):
                ad=[]
                for lane in range(32):
                    q=warp*32+lane; pt=q%PT; s=(q//PT)%G; combo=q//(PT*G); cls=combo&3; py=cls>>1; px=cls&1
                    ty=(pt//TX)*2; tx=(pt%TX)*4
                    ad.append(s*S+(ty+py+r)*SP+tx+px+x)
                tot+=conflicts(ad)
    return tot/(8*15)
best={}
for name,fn,H,pitches in (("e1",conv3,16,(18,19,20)),("e2",conv3,8,(10,11,12,13,14)),("e3",conv3,4,(6,7,8)),("d1",convt,4,(6,7,8,9,10)),("d2",convt,8,(10,11,12,13,14))):
    res=[]
    for SP in pitches:
        for Sm in range(32):
            res.append((fn(H,8192+Sm,SP),SP,Sm))
    res.sort()
    print(name, res[:6])
print("joint B (e2 reads e1p pitch p1, d1 reads e3 pitch p3):")
out=[]
for Sm in range(32):
    for p1 in (10,11,12,13,14):
        for p3 in (6,7,8,9,10):
            out.append((0.17*conv3(8,8192+Sm,p1)+0.31*convt(4,8192+Sm,p3),Sm,p1,p3,conv3(8,8192+Sm,p1),convt(4,8192+Sm,p3)))
out.sort(); print(out[:8])
print("joint A (e1 reads x pitch px, e3 reads e2p pitch p2, d2 reads d1 pitch p4):")
out=[]
for Sm in range(32):
    for p2 in (6,7,8):
        for p4 in (10,11,12,13,14):
            out.append((0.17*conv3(4,8192+Sm,p2)+0.31*convt(8,8192+Sm,p4)+0.02*conv3(16,8192+Sm,18),Sm,p2,p4))
out.sort(); print(out[:8])
