import itertools, sys
def cfg(NP, LD=None, SR=None):
    M=2*NP
    if LD is None:
        pad=(NP//2)%8
        LD=M+(pad if pad else 6)
    if SR is None: SR=8 if (NP>=40 and NP%8==0) else 4
    T=SR*M
    return M,LD,T
def tile_offsets(NP,LD,k,l):
    re=2*k*LD
    oe=[re+l, re+NP+l, re+LD+l, re+LD+NP+l]
    r0=(2*k+1)*LD
    if l<NP-1: oo=[r0+NP+l, r0+l+1, r0+LD+NP+l, r0+LD+l+1]
    elif k<NP-1: oo=[r0+NP+l, NP+k, r0+LD+NP+l, k+1]
    else: oo=[r0+NP+l, NP+l, NP+l, 0]
    return oe,oo
def part_a(NP,a):
    if a<NP: return a,a
    if a<2*NP-1: return a-NP,a-NP+1
    return 0,NP-1
def rest_tile(NP,r):
    k=0; cnt=NP-3
    while r>=cnt:
        r-=cnt; k+=1; cnt=NP-k-2
    return k,k+2+r
def assignment(NP,T,order="row"):
    NT=NP*(NP+1)//2; NA=2*NP; NR=NT-NA; NO=max(T-NA,1); TPO=(NR+NO-1)//NO; TPX=max(TPO,1)
    rest=[rest_tile(NP,r) for r in range(NR)]
    if order=="col":
        rest=sorted(rest,key=lambda kl:(kl[1],kl[0]))
    elif order=="diag":
        rest=sorted(rest,key=lambda kl:(kl[1]-kl[0],kl[0]))
    tiles=[[None]*TPX for _ in range(T)]
    for tid in range(T):
        for qt in range(TPX):
            if tid<NA:
                if qt==0: tiles[tid][qt]=part_a(NP,tid)
            else:
                r=(tid-NA)+qt*NO
                if r<NR: tiles[tid][qt]=rest[r]
    return tiles,TPX
def wavefronts(addrs):
    # 64-bit access: two half-warps of 16 lanes; per half: max over banks (16 banks x 8B) of distinct addresses
    w=0
    for h in range(2):
        lanes=[a for a in addrs[16*h:16*h+16] if a is not None]
        if not lanes: continue
        banks={}
        for a in set(lanes): banks.setdefault(a%16,set()).add(a)
        w+=max(len(v) for v in banks.values())
    return w
def ideal(addrs):
    return sum(1 for h in range(2) if any(a is not None for a in addrs[16*h:16*h+16]))
def simulate(NP,LD=None,order="row"):
    M,LD,T=cfg(NP,LD)
    tiles,TPX=assignment(NP,T,order)
    tot=0; idl=0
    for warp in range(T//32 + (1 if T%32 else 0)):
        for qt in range(TPX):
            for ph in range(2):
                for e in range(4):
                    addrs=[]
                    for lane in range(32):
                        tid=warp*32+lane
                        if tid>=T or tiles[tid][qt] is None: addrs.append(None); continue
                        k,l=tiles[tid][qt]
                        if e==2 and k==l: addrs.append(None); continue   # diag: a10 not accessed
                        oe,oo=tile_offsets(NP,LD,k,l)
                        addrs.append((oo if ph else oe)[e])
                    tot+=wavefronts(addrs); idl+=ideal(addrs)
    return tot,idl,LD
if __name__=="__main__":
    for NP in (12,16,20,24,28,32):
        M=2*NP
        base=simulate(NP)
        res=[]
        for order in ("row","col","diag"):
            for LD in range(M, M+17):
                t,i,_=simulate(NP,LD,order)
                res.append((t,order,LD))
        res.sort()
        print("NP",NP,"current LD",base[2],"wavefronts",base[0],"ideal",base[1],"ratio %.2f"%(base[0]/base[1]),"| best",res[:4])
