import sys
import os; sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from jacobi_banks_rowmajor import cfg, part_a, rest_tile, wavefronts, ideal
def make_layout(NP):
    M,LD,T=cfg(NP)
    NT=NP*(NP+1)//2; NA=2*NP; NR=NT-NA; NO=max(T-NA,1); TPO=(NR+NO-1)//NO; TPX=max(TPO,1)
    PL=TPX*T
    def S(k):
        return 0 if k==0 else (NP-3)+(k-1)*(NP-2)-(k-1)*k//2
    def slot(k,l):
        if k==l: return k
        if l==k+1: return NP+k
        if k==0 and l==NP-1: return 2*NP-1
        r=S(k)+(l-k-2)
        return (r//NO)*T+NA+(r%NO)
    def aidx(r,c):
        if r>c: r,c=c,r
        k,l=r>>1,c>>1
        e=(r&1)*2+(c&1)
        return e*PL+slot(k,l)
    return M,T,NA,NR,NO,TPX,PL,aidx
def tile_rc(NP,k,l,ph):
    m=2*NP
    if not ph: return [(2*k,2*l),(2*k,2*l+1),(2*k+1,2*l),(2*k+1,2*l+1)]
    if l<NP-1: return [(2*k+1,2*l+1),(2*k+1,2*l+2),(2*k+2,2*l+1),(2*k+2,2*l+2)]
    if k<NP-1: return [(2*k+1,m-1),(0,2*k+1),(2*k+2,m-1),(0,2*k+2)]
    return [(m-1,m-1),(0,m-1),(0,m-1),(0,0)]
def simulate(NP):
    M,T,NA,NR,NO,TPX,PL,aidx=make_layout(NP)
    tiles=[[None]*TPX for _ in range(T)]
    for tid in range(T):
        for qt in range(TPX):
            if tid<NA:
                if qt==0: tiles[tid][qt]=part_a(NP,tid)
            else:
                r=(tid-NA)+qt*NO
                if r<NR: tiles[tid][qt]=rest_tile(NP,r)
    # sanity: aidx injective on upper triangle
    seen={}
    for r in range(M):
        for c in range(r,M):
            a=aidx(r,c); assert a not in seen,(r,c,seen[a]); seen[a]=(r,c)
    assert max(seen)<4*PL
    tot=idl=0
    for warp in range((T+31)//32):
        for qt in range(TPX):
            for ph in range(2):
                for e in range(4):
                    addrs=[]
                    for lane in range(32):
                        tid=warp*32+lane
                        if tid>=T or tiles[tid][qt] is None: addrs.append(None); continue
                        k,l=tiles[tid][qt]
                        if e==2 and k==l: addrs.append(None); continue
                        r,c=tile_rc(NP,k,l,ph)[e]
                        addrs.append(aidx(r,c))
                    tot+=wavefronts(addrs); idl+=ideal(addrs)
    return tot,idl,4*PL,M
for NP in (4,8,12,16,20,24,28,32,40,48,56,64):
    t,i,sz,M=simulate(NP)
    print("NP",NP,"wavefronts",t,"ideal",i,"ratio %.2f"%(t/i),"storage doubles",sz,"(old M*LD ~",M*(M+6),")")
