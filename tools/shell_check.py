import random
def rsort(x, idx):
    n=len(x); h=1
    while h <= n//9: h=3*h+1
    while h>0:
        for i in range(h,n):
            v=x[i]; iv=idx[i]; j=i
            while j>=h and x[j-h]>v:
                x[j]=x[j-h]; idx[j]=idx[j-h]; j-=h
            x[j]=v; idx[j]=iv
        h//=3
def emu(x, idx, MC):
    n=len(x); NEG=float('-inf')
    d=[NEG]*(MC-n)+list(x); ii=[0]*(MC-n)+list(idx)
    def cswap(a,b):
        if a>=0 and d[a]>d[b]:
            d[a],d[b]=d[b],d[a]; ii[a],ii[b]=ii[b],ii[a]; return True
        return False
    if n>8:
        moving=True
        c=MC-1
        while c>=4:
            sw = moving and d[c-4]>d[c]
            if sw:
                d[c-4],d[c]=d[c],d[c-4]; ii[c-4],ii[c]=ii[c],ii[c-4]
            moving=sw
            c-=4
        c=(MC-1)%4
        while c<MC:
            cswap(c-1,c); cswap(c-2,c-1); cswap(c-3,c-2)
            c+=4
    else:
        c=MC-1
        while c>=1 and c>=MC-8:
            cswap(c-1,c); c-=1
    return d[MC-n:], ii[MC-n:]
random.seed(1)
bad=0
for trial in range(200000):
    n=random.randint(1,32)
    MC=[m for m in (8,16,24,32) if m>=n][0]
    vals=sorted(random.choice([random.randint(0,6), random.random()*6, float('inf')]) for _ in range(n))
    idx=list(range(100,100+n))
    # new candidate smaller than worst
    if vals[-1]==0: continue
    v=random.choice([random.randint(0,6), random.random()*6])
    if not v<vals[-1]: continue
    x=list(vals); x[-1]=v; ix=list(idx); ix[-1]=999
    a=list(x); ai=list(ix); rsort(a,ai)
    b,bi=emu(x,ix,MC)
    if a!=b or ai!=bi:
        bad+=1
        if bad<5: print("MISMATCH",n,vals,v,a,ai,b,bi)
print("bad",bad)
