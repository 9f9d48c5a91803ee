"""Functional emulation of wgrad_tc2.cu's tile/coordinate logic (not of the hardware): TMA boxes = zero-filled numpy slices,
a pair MMA = A_pair^T @ B_pair, epilogue addressing as in the kernel.  Compared with torch's conv2d_weight."""
import numpy as np, torch, math
def flat_rows(N,H,W): return N*(H+1)*(W+1)+ (W+1) + 1   # generous alloc
def to_flat(x):  # (N,C,H,W) -> [rows][C] padded-flat
    N,C,H,W = x.shape; Wp, Hp = W+1, H+1
    rows = N*Hp*Wp
    out = np.zeros((rows + 64, C), np.float32)
    for n in range(N):
        for h in range(H):
            r0 = n*Hp*Wp + (h+1)*Wp + 1
            out[r0:r0+W] = x[n,:,h,:].T
    return out, rows
def box(mat, c0, r0, nrows=64, ncols=64):   # TMA 2-D box with zero fill (coords may be negative / past the end)
    out = np.zeros((nrows, ncols), np.float32)
    for i in range(nrows):
        r = r0+i
        if 0 <= r < mat.shape[0]:
            lo, hi = max(c0,0), min(c0+ncols, mat.shape[1])
            if hi > lo: out[i, lo-c0:hi-c0] = mat[r, lo:hi]
    return out
def emulate(N,H,W,cin,cout,k, pairs=74):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N,cin,H,W,generator=g); dy = torch.randn(N,cout,H,W,generator=g)
    ref = torch.nn.grad.conv2d_weight(x.double(), (cout,cin,k,k), dy.double(), padding=(k-1)//2).float().numpy()
    X, m_rows = to_flat(x.numpy()); DY,_ = to_flat(dy.numpy())
    taps = k*k; Wp = W+1
    shifts = [((t//3)-1)*Wp + ((t%3)-1) if taps==9 else 0 for t in range(9)]
    kblocks_total = (m_rows+63)//64
    co_pairs, ci_tiles = cout//256, cin//256
    base_items = co_pairs*ci_tiles*taps
    max_splits = max(kblocks_total//32,1); best=1; best_eff=0
    for s in range(1, min(max_splits,8)+1):
        items=base_items*s; waves=(items+pairs-1)//pairs; eff=items/(waves*pairs)-0.02*(s-1)
        if eff>best_eff+1e-9: best_eff=eff; best=s
    splits=best; kb_per_split=(kblocks_total+splits-1)//splits
    dw = np.zeros((cout, taps, cin), np.float32)
    items = base_items*splits
    for it in range(items):
        t=it; co_p=t%co_pairs; t//=co_pairs; ci_t=t%ci_tiles; t//=ci_tiles; tap=t%taps; t//=taps
        kb0=t*kb_per_split; kb1=min(kb0+kb_per_split, kblocks_total)
        D = np.zeros((256,256), np.float32)
        for kb in range(kb0,kb1):
            row=kb*64
            A=[]; B=[]
            for rank in (0,1):
                co0=co_p*256+rank*128; ci0=ci_t*256+rank*128
                A.append(np.concatenate([box(DY,co0,row), box(DY,co0+64,row)],1))          # [64 rows][128 co]
                B.append(np.concatenate([box(X,ci0,row+shifts[tap]), box(X,ci0+64,row+shifts[tap])],1))
            Ap=np.concatenate(A,1); Bp=np.concatenate(B,1)     # M = rank0 co | rank1 co ; N = rank0 ci | rank1 ci
            D += Ap.T @ Bp
        if kb1>kb0:
            for rank in (0,1):
                for lane_all in range(128):
                    co = co_p*256 + rank*128 + lane_all
                    dw[co, tap, ci_t*256:(ci_t+1)*256] += D[rank*128+lane_all]
    out = dw.reshape(cout,k,k,cin).transpose(0,3,1,2)
    err = np.abs(out-ref).max()/np.abs(ref).max()
    print((N,H,W,cin,cout,k), 'splits',splits,'items',items,'rel err %.2e'%err)
    return err
assert emulate(2,13,13,256,256,3) < 1e-5
assert emulate(1,13,13,512,256,1) < 1e-5
assert emulate(3,7,9,256,512,3, pairs=4) < 1e-5
print('wgrad_tc2 tile logic ok')
