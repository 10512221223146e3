"""Per-layer model of the 608x608 batch-32 conv stack: for each of the 75 layers the tile configuration tc_make_plan /
halo_make_plan choose, the bytes that cross L2 -> SM (A and B operand loads, residual loads, stores), and three lower
bounds -- tensor (sustained peak), HBM (algorithmic bytes) and L2 -> SM (at ~6.3 KB/clk, B300_MICROARCH.md "LTS throughput
cap") -- next to the per-launch times of a committed ncu launch list (profiles/r02_launches.csv; L2_MODEL_LAUNCHES=<file> selects another).  Layers whose L2 -> SM
time exceeds both other bounds are flagged: that is how the 64->128 stride-2 layer was found to be bound by re-fetching
its 147 KB weight slab for every tile.  Run here (no GPU needed): python tools/l2_model.py
"""
import csv
B=32; H=608
layers=[]
def add(name,cin,cout,ks,s,hin,res=False,head=False,up=False): layers.append(dict(name=name,cin=cin,cout=cout,ks=ks,s=s,hin=hin,res=res,head=head,up=up))
add('stem',3,32,3,1,608)
ch=32;h=608
blocks=[1,2,8,8,4]
for st in range(5):
    add(f'down{st}',ch,ch*2,3,2,h); ch*=2; h//=2
    for j in range(blocks[st]):
        add(f's{st}r{j}c1',ch,ch//2,1,1,h); add(f's{st}r{j}c2',ch//2,ch,3,1,h,res=True)
def predet(nm,nin,nout,h):
    for i in range(3):
        add(f'{nm}.{2*i}',nin,nout,1,1,h); add(f'{nm}.{2*i+1}',nout,nout*2,3,1,h); nin=nout*2
    add(f'{nm}.head',nin,255,1,1,h,head=True)
predet('pd1',1024,512,19); add('up1',512,256,1,1,19,up=True)
predet('pd2',768,256,38); add('up2',256,128,1,1,38,up=True)
predet('pd3',384,128,76)
import os
rows=list(csv.reader(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', os.environ.get('L2_MODEL_LAUNCHES', 'r02_launches.csv')))))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
t=[float(r[-1])/1000 for r in rows[hi+1:hi+76]]
nsm=148
tot_meas=0;tot_l2=0;tot_hbm=0;tot_tc=0
print(f"{'#':>2} {'name':10} {'cin':>4} {'cout':>4} k s {'ho':>3} BN c2 br | {'meas':>6} {'tc':>6} {'hbm':>6} {'l2':>6} us | l2GB")
for i,L in enumerate(layers):
    ho=L['hin']//L['s']; M=B*ho*ho; cin=L['cin']; cout_pad=(L['cout']+15)//16*16
    K=L['ks']**2*cin
    flops=2*M*L['cout']*K
    esz=4 if L['head'] else 2
    hbm=(B*L['hin']**2*cin*(4 if i==0 else 2)+M*cout_pad*esz*(4 if L['up'] else 1)+(M*L['cout']*2 if L['res'] else 0)+K*cout_pad*2)
    if i==0 or cin==3:
        l2=hbm; bn=32; c2=0; br=1
    else:
        swz=64 if cin==32 else 128; bke=swz//2
        nkb=L['ks']**2*(cin//bke)
        m_tiles=(M+127)//128
        best=None
        for bn_ in range(min(cout_pad,256),15,-16):
            if cout_pad%bn_: continue
            if bn_<64 and bn_!=cout_pad: break
            tiles=m_tiles*(cout_pad//bn_); waves=(tiles+nsm-1)//nsm; cost=waves*(128+bn_)
            if best is None or cost<best[1]: best=(bn_,cost)
        bn=best[0]; n_tiles=cout_pad//bn
        c2=int(swz==128 and bn==256 and not L['up'] and m_tiles>=4)
        ncta=2 if c2 else 1
        mt=(m_tiles+1)//2 if c2 else m_tiles
        grid=2*min(mt*n_tiles,nsm//2) if c2 else min(mt*n_tiles,nsm)
        bslot=((bn//ncta)*swz+1023)//1024*1024
        br=int((not c2) and bslot*nkb<=96*1024 and grid%n_tiles==0 and mt>2*nsm)
        halo = L['ks']==3 and ((cin in (32,64) and L['s']==1) or (cin==32 and L['s']==2)) and L['cout'] in (64,128) and ho%38==0
        if halo:
            # patch once per tile of 114 outputs: 5x40 (s1) or 4 planes 4x40 (s2) pixels x cin x2B ; weights resident
            tiles=M/114
            a_bytes=tiles*(200 if L['s']==1 else 640)*cin*2
            l2=a_bytes+grid*cout_pad*K*2
            bn=cout_pad; br=1; c2=0; bn=-bn
        else:
            tiles=m_tiles*n_tiles  # 128-row tiles x n
            a_per=128*swz*nkb
            b_per=bn*swz*nkb
            if c2:
                # per pair tile: A 2x128 rows, B bn rows (half each)
                l2=mt*n_tiles*(2*a_per+b_per)
            else:
                l2=tiles*a_per+(grid*b_per if br else tiles*b_per)
        l2+= M*cout_pad*esz*(4 if L['up'] else 1) + (M*L['cout']*2 if L['res'] else 0)   # stores + residual loads also cross the xbar
    t_tc=flops/1386.8e12*1e6; t_hbm=hbm/6445e9*1e6; t_l2=l2/(6300*1.9e9)*1e6
    tot_meas+=t[i]; tot_tc+=t_tc; tot_hbm+=t_hbm; tot_l2+=t_l2
    flag=' <-- L2' if t_l2>max(t_tc,t_hbm)*1.05 else ''
    print(f"{i:2d} {L['name']:10} {cin:4d} {L['cout']:4d} {L['ks']} {L['s']} {ho:3d} {bn:3d} {c2:2d} {br:2d} | {t[i]:6.1f} {t_tc:6.1f} {t_hbm:6.1f} {t_l2:6.1f}    | {l2/1e9:5.2f}{flag}")
print('sum meas',tot_meas,'sum max bound',)
