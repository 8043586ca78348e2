// TEST INFRASTRUCTURE ONLY (design model, not shipped, not called by the product).
// CPU model of the order-exact PARALLEL formulation of UpdateESDF used by fiesta_b200/csrc/fb_exact.cu (persistent
// kernel k_x_relax: one FIFO generation after the other on the device; per generation a behaviour fixpoint over
// work lists), checked against the sequential oracle in the same process:
//   gcc -O2 -ffp-contract=off -o /tmp/exact_model oracle/exact_model.c -lm && /tmp/exact_model 32 0.7 6 1500 1 [small]
// prints, per update, the reference's expansion count, ours, and the number of distance / closest-obstacle mismatches
// (all 0).  Every "for each ..." loop of a phase is data-parallel in the kernel; here its items run one after the other
// in a RANDOM order (in-place updates), which is one legal interleaving of the kernel's threads.
//
// Per generation (E = entries in queue order; element i, direction k acts at timestamp 32*i+k, its pull at 32*i+24):
//   MB[v]   packed word of the live entry at voxel v: {queue position, behaviour (dead/pull/push), code}.
//   state(v,T) = snapshot, or the lexicographic minimum (distance, timestamp) over the offers with timestamp < T of the
//             <= 25 elements that can write v (gather()).  BIG generations cache per target a summary
//             {first improving timestamp, best timestamp, best code, snapshot code} (summarize()); a query only
//             gathers when first < T <= best.
//   round 1 evaluates every element; a flip (behaviour change) lists every LATER element whose inputs it can touch
//             (the 129 offsets a+b, a,b in {0} u dirs) for the next round and, in BIG mode, is itself listed so that
//             the summaries of its targets are recomputed during the next round.  Rounds after the first evaluate by
//             gathering (summaries may be stale there) unless the work list is dense, in which case all summaries are
//             recomputed first.  The fixpoint is reached by a round without flips: it has read final words only.
//   commit: element i owns slot k iff the final state of its target k carries timestamp 32*i+k; the owned slots in
//             timestamp order are the next generation (per-element masks + exclusive scan).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "esdf_oracle.c"
typedef unsigned int u32; typedef unsigned long long u64;
#define NONE 0xffffffffu
#define CU 0u   /* unknown */
#define CI 1u   /* inf */
#define K_DEAD 0ull
#define K_PULL 1ull
#define K_PUSH 2ull
#define MB_NONE 0xffffffffffffffffull
static int GX,GY,GZ; static long N;
static u32 *C;            // codes
static u64 *MB;           // packed live-entry words
static u64 *LS;           // link sequence (time of last relink)
static u64 tclock=1;
static omap* O;
static long SMALL=64;     // generations up to this size run in SMALL mode (no summaries)
static inline u32 pack(int x,int y,int z){return ((u32)(x+1)<<20)|((u32)y<<10)|(u32)z;}
static inline void unpack(u32 c,int*x,int*y,int*z){*x=(int)(c>>20)-1;*y=(c>>10)&1023;*z=c&1023;}
static inline long vi(int x,int y,int z){return ((long)x*GY+y)*GZ+z;}
static inline int ing(int x,int y,int z){return x>=0&&y>=0&&z>=0&&x<GX&&y<GY&&z<GZ;}
static inline int inb(int x,int y,int z){return x>=O->min_vec[0]&&x<=O->max_vec[0]&&y>=O->min_vec[1]&&y<=O->max_vec[1]&&z>=O->min_vec[2]&&z<=O->max_vec[2];}
static inline u32 d2(int x,int y,int z,u32 c){int ox,oy,oz;unpack(c,&ox,&oy,&oz);ox-=x;oy-=y;oz-=z;return (u32)(ox*ox+oy*oy+oz*oz);}
static inline int existc(u32 c){int x,y,z;unpack(c,&x,&y,&z);return O->occ[vi(x,y,z)]>O->l_occ;}
#define DINF 0xffffffffu
static inline u32 dcode(int x,int y,int z,u32 c){ return c<2?DINF:d2(x,y,z,c); }
static inline void vxyz(long v,int*x,int*y,int*z){*x=v/(GY*GZ);*y=(v/GZ)%GY;*z=v%GZ;}
static inline u64 mbw(u32 i,u64 kind,u32 code){return ((u64)i<<33)|(kind<<31)|(u64)(code&0x7fffffffu);}
static inline u32 mb_idx(u64 w){return (u32)(w>>33);} static inline u64 mb_kind(u64 w){return (w>>31)&3ull;} static inline u32 mb_code(u64 w){return (u32)(w&0x7fffffffull);}

typedef struct {u32 d,c,ts;} st_t;
typedef struct {u32 first,best_ts,best_c,snap_c;} sum_t;
static sum_t *SUM; static u32 *SUMg; static u32 gen_id=0;
static long n_gather=0,n_query=0;
// State of voxel v as seen at time T (exclusive): gather over the <= 25 writers.
static st_t gather(int x,int y,int z,u32 T,u32*first){
  st_t s; long v=vi(x,y,z); s.c=C[v]; s.d=dcode(x,y,z,s.c); s.ts=NONE; u32 d0=s.d; if(first)*first=NONE; n_gather++;
  if(s.c==CU) return s;          // unknown voxels never accept (:382)
  if(inb(x,y,z))                 // pushes only go to in-box voxels (:378); an entry's own pull is not range-checked (:349-367)
  for(int k=0;k<24;k++){ int qx=x-DIRS[k][0],qy=y-DIRS[k][1],qz=z-DIRS[k][2]; if(!ing(qx,qy,qz)) continue;
    u64 w=MB[vi(qx,qy,qz)]; if(w==MB_NONE||mb_kind(w)!=K_PUSH) continue; u32 ts=mb_idx(w)*32+k; u32 c=mb_code(w); u32 d=d2(x,y,z,c);
    if(d<d0){ if(first&&ts<*first)*first=ts; if(ts<T&&(d<s.d||(d==s.d&&ts<s.ts))){s.d=d;s.c=c;s.ts=ts;} } }
  u64 w=MB[v]; if(w!=MB_NONE&&mb_kind(w)==K_PULL){ u32 ts=mb_idx(w)*32+24; u32 c=mb_code(w); u32 d=d2(x,y,z,c);
    if(d<d0){ if(first&&ts<*first)*first=ts; if(ts<T&&(d<s.d||(d==s.d&&ts<s.ts))){s.d=d;s.c=c;s.ts=ts;} } }
  return s;
}
static void summarize(int x,int y,int z){ u32 first; st_t f=gather(x,y,z,NONE,&first); sum_t u; u.first=first; u.best_ts=f.ts; u.best_c=f.c; u.snap_c=C[vi(x,y,z)]; SUM[vi(x,y,z)]=u; }
static st_t state_sum(int x,int y,int z,u32 T){ n_query++;
  sum_t u=SUM[vi(x,y,z)]; st_t s;
  if(u.first==NONE||T<=u.first){ s.c=u.snap_c; s.d=dcode(x,y,z,s.c); s.ts=NONE; return s; }
  if(T>u.best_ts){ s.c=u.best_c; s.d=dcode(x,y,z,s.c); s.ts=u.best_ts; return s; }
  return gather(x,y,z,T,NULL);
}
static u32 *E[2]; static long nE; static int cur=0;
// behaviour of element i given the current words (ESDFMap.cpp:345-373)
static u64 eval(long i,int use_sum){
  long p=E[cur][i]; int x,y,z; vxyz(p,&x,&y,&z); u32 T0=(u32)i*32;
  st_t s=use_sum?state_sum(x,y,z,T0):gather(x,y,z,T0,NULL);
  u32 d0=dcode(x,y,z,C[p]);
  if(s.d!=d0) return mbw((u32)i,K_DEAD,0);
  u32 curd=s.d,curc=s.c; int ch=0;
  for(int k=0;k<24;k++){ int nx=x+DIRS[k][0],ny=y+DIRS[k][1],nz=z+DIRS[k][2]; if(!ing(nx,ny,nz)||!inb(nx,ny,nz)) continue;
    st_t sn=use_sum?state_sum(nx,ny,nz,T0):gather(nx,ny,nz,T0,NULL); if(sn.c<2) continue; u32 t=d2(x,y,z,sn.c); if(curd>t){curd=t;curc=sn.c;ch=1;} }
  return ch?mbw((u32)i,K_PULL,curc):mbw((u32)i,K_PUSH,s.c);
}
// the 129 distinct offsets a+b
static int OFF[1024][3]; static int nOFF=0;
static void build_offsets(void){ for(int a=-1;a<24;a++)for(int b=-1;b<24;b++){ int o[3]; for(int q=0;q<3;q++)o[q]=(a<0?0:DIRS[a][q])+(b<0?0:DIRS[b][q]);
  int dup=0; for(int j=0;j<nOFF;j++) if(OFF[j][0]==o[0]&&OFF[j][1]==o[1]&&OFF[j][2]==o[2])dup=1; if(!dup){OFF[nOFF][0]=o[0];OFF[nOFF][1]=o[1];OFF[nOFF][2]=o[2];nOFF++;} } }
static u32 *W[3]; static long nW[3]; static u32 *wstamp; static u32 wclock=0;
static void shuffle(u32*a,long n){ for(long i=n-1;i>0;i--){ long j=rand()%(i+1); u32 t=a[i];a[i]=a[j];a[j]=t; } }
static long expansions; static long totrounds=0, totgens=0, maxrounds=0, totevals=0, totvisits=0, totpasses=0;
static u32 *emask, *ecode;
// ---- BIG generations: tile mode.  The grid is cut into 8^3 tiles; a visit stages the tile + a 4-voxel halo (16^3 words and
// codes), finds the tile's entries in the staged words, iterates them to a LOCAL fixpoint (dirty bits; flips are written to
// the staged copy and to the global words), computes the slot masks of all its entries from the staged copy, and lists the
// neighbour tiles within reach (4 voxels) of a flip for the next round.  A round without flips has staged final words only.
static int TX,TY,TZ; static u32 *tstamp,*dstamp; static u32 *AT[2]; static long nAT[2]; static u32 *DT[3]; static long nDT[3];
static int MAXPASS=4;
typedef struct { u64 mb[4096]; u32 cb[4096]; int ox,oy,oz; } stage_t;
static inline int sidx(int lx,int ly,int lz){return (lx*16+ly)*16+lz;}
static void stage(stage_t*S,int t){ int tz=t%TZ,ty=(t/TZ)%TY,tx=t/(TZ*TY); S->ox=8*tx-4;S->oy=8*ty-4;S->oz=8*tz-4;
  for(int lx=0;lx<16;lx++)for(int ly=0;ly<16;ly++)for(int lz=0;lz<16;lz++){ int x=S->ox+lx,y=S->oy+ly,z=S->oz+lz; int k=sidx(lx,ly,lz);
    if(ing(x,y,z)){S->mb[k]=MB[vi(x,y,z)];S->cb[k]=C[vi(x,y,z)];} else {S->mb[k]=MB_NONE;S->cb[k]=CU;} } }
static st_t gather_l(const stage_t*S,int x,int y,int z,u32 T){   // gather() on the staged copy (global coordinates)
  st_t s; int lx=x-S->ox,ly=y-S->oy,lz=z-S->oz; s.c=S->cb[sidx(lx,ly,lz)]; s.d=dcode(x,y,z,s.c); s.ts=NONE; u32 d0=s.d;
  if(s.c==CU) return s;
  if(inb(x,y,z))
  for(int k=0;k<24;k++){ int qx=lx-DIRS[k][0],qy=ly-DIRS[k][1],qz=lz-DIRS[k][2]; if(qx<0||qy<0||qz<0||qx>15||qy>15||qz>15){printf("stage reach\n");exit(1);}
    u64 w=S->mb[sidx(qx,qy,qz)]; if(w==MB_NONE||mb_kind(w)!=K_PUSH) continue; u32 ts=mb_idx(w)*32+k; u32 c=mb_code(w); u32 d=d2(x,y,z,c);
    if(d<d0&&ts<T&&(d<s.d||(d==s.d&&ts<s.ts))){s.d=d;s.c=c;s.ts=ts;} }
  u64 w=S->mb[sidx(lx,ly,lz)]; if(w!=MB_NONE&&mb_kind(w)==K_PULL){ u32 ts=mb_idx(w)*32+24; u32 c=mb_code(w); u32 d=d2(x,y,z,c);
    if(d<d0&&ts<T&&(d<s.d||(d==s.d&&ts<s.ts))){s.d=d;s.c=c;s.ts=ts;} }
  return s;
}
static u64 eval_l(const stage_t*S,u32 i,int x,int y,int z){
  u32 T0=i*32; st_t s=gather_l(S,x,y,z,T0); u32 d0=dcode(x,y,z,S->cb[sidx(x-S->ox,y-S->oy,z-S->oz)]);
  if(s.d!=d0) return mbw(i,K_DEAD,0);
  u32 curd=s.d,curc=s.c; int ch=0;
  for(int k=0;k<24;k++){ int nx=x+DIRS[k][0],ny=y+DIRS[k][1],nz=z+DIRS[k][2]; if(!ing(nx,ny,nz)||!inb(nx,ny,nz)) continue;
    st_t sn=gather_l(S,nx,ny,nz,T0); if(sn.c<2) continue; u32 t=d2(x,y,z,sn.c); if(curd>t){curd=t;curc=sn.c;ch=1;} }
  return ch?mbw(i,K_PULL,curc):mbw(i,K_PUSH,s.c);
}
static void mark_tile(u32 t,int out){ if(dstamp[t]!=wclock){ dstamp[t]=wclock; DT[out][nDT[out]++]=t; } }
// one visit; returns the number of flips
static long visit(stage_t*S,int t,int out){
  int tz=t%TZ,ty=(t/TZ)%TY,tx=t/(TZ*TY); totvisits++;
  static u32 el[512]; int nel=0; unsigned char dirty[512],next[512]; memset(dirty,0,512);
  for(int j=0;j<512;j++){ int lx=4+(j>>6),ly=4+((j>>3)&7),lz=4+(j&7); u64 w=S->mb[sidx(lx,ly,lz)]; if(w!=MB_NONE){ el[nel++]=j; dirty[j]=1; } }
  long flips=0; u32 nbmask=0; int left=nel;
  for(int pass=0;pass<MAXPASS&&left;pass++){ totpasses++; memset(next,0,512); left=0;
    u32 ord[512]; int no=0; for(int q=0;q<nel;q++) if(dirty[el[q]]) ord[no++]=el[q]; shuffle(ord,no);
    for(int q=0;q<no;q++){ int j=ord[q]; int lx=4+(j>>6),ly=4+((j>>3)&7),lz=4+(j&7); int x=S->ox+lx,y=S->oy+ly,z=S->oz+lz; u64 w=S->mb[sidx(lx,ly,lz)]; u32 i=mb_idx(w); totevals++;
      u64 nb=eval_l(S,i,x,y,z);
      if(nb!=w){ S->mb[sidx(lx,ly,lz)]=nb; MB[vi(x,y,z)]=nb; flips++;
        for(int o=0;o<nOFF;o++){ int qx=lx+OFF[o][0],qy=ly+OFF[o][1],qz=lz+OFF[o][2]; if(qx<4||qy<4||qz<4||qx>11||qy>11||qz>11) continue;
          u64 w2=S->mb[sidx(qx,qy,qz)]; if(w2!=MB_NONE&&mb_idx(w2)>i){ int j2=((qx-4)<<6)|((qy-4)<<3)|(qz-4); if(!next[j2]){next[j2]=1;left++;} } }
        int dx=(lx-4)<=3?-1:1,dy=(ly-4)<=3?-1:1,dz=(lz-4)<=3?-1:1;
        for(int sx=0;sx<2;sx++)for(int sy=0;sy<2;sy++)for(int sz=0;sz<2;sz++){ if(!(sx|sy|sz)) continue; int ax=1+sx*dx,ay=1+sy*dy,az=1+sz*dz; nbmask|=1u<<((ax*3+ay)*3+az); } } }
    memcpy(dirty,next,512); }
  if(left) mark_tile((u32)t,out);                      // pass budget exhausted: come back next round
  for(int q=0;q<nel;q++){ int j=el[q]; int lx=4+(j>>6),ly=4+((j>>3)&7),lz=4+(j&7); int x=S->ox+lx,y=S->oy+ly,z=S->oz+lz; u64 b=S->mb[sidx(lx,ly,lz)]; u32 i=mb_idx(b); u32 m=0;
    if(mb_kind(b)==K_PUSH){ for(int k=0;k<24;k++){ int nx=x+DIRS[k][0],ny=y+DIRS[k][1],nz=z+DIRS[k][2]; if(!ing(nx,ny,nz)||!inb(nx,ny,nz)) continue; st_t f=gather_l(S,nx,ny,nz,NONE); if(f.ts==i*32+k) m|=1u<<k; } }
    else if(mb_kind(b)==K_PULL){ st_t f=gather_l(S,x,y,z,NONE); if(f.ts==i*32+24) m|=1u<<24; }
    emask[i]=m; }
  for(int a=0;a<27;a++) if(nbmask>>a&1){ int ax=a/9-1,ay=(a/3)%3-1,az=a%3-1; int nx=tx+ax,ny=ty+ay,nz=tz+az; if(nx<0||ny<0||nz<0||nx>=TX||ny>=TY||nz>=TZ) continue; u32 nt=(u32)((nx*TY+ny)*TZ+nz);
      if(tstamp[nt]==gen_id) mark_tile(nt,out); }
  return flips;
}
static void begin_entry(long r,long v,u32 c,int big,int par){   // new entry r of a generation at voxel v with code c
  MB[v]=mbw((u32)r,K_PUSH,c);
  if(big){ int x,y,z; vxyz(v,&x,&y,&z); u32 t=(u32)(((x>>3)*TY+(y>>3))*TZ+(z>>3)); if(tstamp[t]!=gen_id){ tstamp[t]=gen_id; AT[par][nAT[par]++]=t; } }
}
static void relax(void){
  int big=nE>SMALL; gen_id++; nAT[gen_id&1]=0;
  for(long i=0;i<nE;i++) begin_entry(i,E[cur][i],C[E[cur][i]],big,gen_id&1);
  while(nE){
    totgens++;
    int rounds=0; nW[0]=nW[1]=nW[2]=0; nDT[0]=nDT[1]=nDT[2]=0;
    for(int r=1;;r++){
      int in=r%3,out=(r+1)%3; nW[(r+2)%3]=0; nDT[(r+2)%3]=0; wclock++;
      if(big){
        long nt=r==1?nAT[gen_id&1]:nDT[in]; u32*tl=r==1?AT[gen_id&1]:DT[in];
        if(r>1&&nt==0) break;
        rounds++; shuffle(tl,nt);
        int pre=rand()&1;                                  // stage every tile at the start of the round (stalest legal view) or when it is visited
        stage_t*SS=pre?malloc(sizeof(stage_t)*nt):malloc(sizeof(stage_t)); if(pre) for(long q=0;q<nt;q++) stage(&SS[q],tl[q]);
        for(long q=0;q<nt;q++){ if(!pre) stage(&SS[0],tl[q]); visit(pre?&SS[q]:&SS[0],tl[q],out); }
        free(SS);
      } else {
        long nw; u32*wl=W[in];
        if(r==1){ nw=nE; for(long i=0;i<nE;i++)wl[i]=(u32)i; } else nw=nW[in];
        if(r>1&&nw==0) break;
        rounds++; shuffle(wl,nw);
        for(long q=0;q<nw;q++){ long i=wl[q]; totevals++;
          u64 nb=eval(i,0); long p=E[cur][i];
          if(nb!=MB[p]){ MB[p]=nb;
            int x,y,z; vxyz(p,&x,&y,&z);
            for(int o=0;o<nOFF;o++){ int nx=x+OFF[o][0],ny=y+OFF[o][1],nz=z+OFF[o][2]; if(!ing(nx,ny,nz)) continue; u64 w=MB[vi(nx,ny,nz)]; if(w==MB_NONE) continue; u32 j=mb_idx(w);
              if(j>(u32)i && wstamp[j]!=wclock){ wstamp[j]=wclock; W[out][nW[out]++]=j; } } } }
      }
      if(rounds>100000){printf("no convergence\n");exit(1);} }
    totrounds+=rounds; if(rounds>maxrounds)maxrounds=rounds;
    // count phase: SMALL generations compute their slot masks here (by gathering); every entry's code and liveness are taken from its final word
    long total=0;
    for(long i=0;i<nE;i++){ long p=E[cur][i]; u64 b=MB[p]; int x,y,z; vxyz(p,&x,&y,&z);
      if(mb_kind(b)!=K_DEAD) expansions++;
      ecode[i]=mb_code(b);
      if(!big){ u32 m=0;
        if(mb_kind(b)==K_PUSH){ for(int k=0;k<24;k++){ int nx=x+DIRS[k][0],ny=y+DIRS[k][1],nz=z+DIRS[k][2]; if(!ing(nx,ny,nz)||!inb(nx,ny,nz)) continue; st_t f=gather(nx,ny,nz,NONE,NULL); if(f.ts==(u32)i*32+k) m|=1u<<k; } }
        else if(mb_kind(b)==K_PULL){ st_t f=gather(x,y,z,NONE,NULL); if(f.ts==(u32)i*32+24) m|=1u<<24; }
        emask[i]=m; }
      total+=__builtin_popcount(emask[i]); }
    for(long i=0;i<nE;i++) MB[E[cur][i]]=MB_NONE;    // (kernel: retired by compare-and-swap in the apply phase)
    // apply: exclusive scan of the mask popcounts -> positions; an owned slot carries the code its owner offered
    int big2=total>SMALL; gen_id++; nAT[gen_id&1]=0; long r=0;
    for(long i=0;i<nE;i++){ long p=E[cur][i]; int x,y,z; vxyz(p,&x,&y,&z); u32 m=emask[i];
      for(int k=0;k<25;k++) if(m>>k&1){ int nx=k<24?x+DIRS[k][0]:x,ny=k<24?y+DIRS[k][1]:y,nz=k<24?z+DIRS[k][2]:z; long v=vi(nx,ny,nz);
          u32 c=ecode[i]; C[v]=c; LS[v]=tclock+(u64)i*32+k; E[cur^1][r]=(u32)v; begin_entry(r,v,c,big2,gen_id&1); r++; } }
    tclock+=(u64)nE*32+1; nE=r; cur^=1; big=big2;
  }
}
typedef struct {u64 k1,k2; u32 v;} dep_t;
static int cmpdep(const void*a,const void*b){const dep_t*x=a,*y=b; if(x->k1!=y->k1) return x->k1<y->k1?-1:1; if(x->k2!=y->k2) return x->k2>y->k2?-1:1; return 0;}
int main(int argc,char**argv){
  int G=argc>1?atoi(argv[1]):24; double obs=argc>2?atof(argv[2]):0.7; int rounds=argc>3?atoi(argv[3]):6; int nops=argc>4?atoi(argv[4]):600; srand(argc>5?atoi(argv[5]):1);
  if(argc>6) SMALL=atol(argv[6]);
  int local=argc>7?atoi(argv[7]):0;   // 1: shrink the update box for the later rounds (SetUpdateRange)
  double org[3]={0,0,0},sz[3]={G*0.1-0.05,G*0.1-0.05,G*0.1-0.05}; O=fiesta_oracle_create(org,0.1,sz); fiesta_oracle_set_parameters(O,0.97,0.03,0.30,0.90,0.80);
  GX=O->gs[0];GY=O->gs[1];GZ=O->gs[2];N=(long)GX*GY*GZ; C=calloc(N,4); MB=malloc(N*8); LS=calloc(N,8); for(long i=0;i<N;i++)MB[i]=MB_NONE;
  if(argc>8) MAXPASS=atoi(argv[8]);
  emask=malloc(4*N); ecode=malloc(4*N);
  E[0]=malloc(4*N);E[1]=malloc(4*N); for(int q=0;q<3;q++){W[q]=malloc(4*N);} wstamp=calloc(N,4);
  TX=(GX+7)/8;TY=(GY+7)/8;TZ=(GZ+7)/8; long NTL=(long)TX*TY*TZ; tstamp=calloc(NTL,4);dstamp=calloc(NTL,4); for(int q=0;q<2;q++)AT[q]=malloc(4*NTL); for(int q=0;q<3;q++)DT[q]=malloc(4*NTL);
  build_offsets(); if(nOFF!=129){printf("offsets %d\n",nOFF);return 1;}
  long bad=0;
  for(int r=0;r<rounds;r++){
    if(local&&r==2){ double mn[3]={0.4,0.5,0.3},mx[3]={G*0.1-0.6,G*0.1-0.4,G*0.1-0.5}; fiesta_oracle_set_update_range(O,mn,mx,1); }
    int nev=r==0?(int)(N*obs):nops;
    for(int e=0;e<nev;e++){int v[3]={rand()%GX,rand()%GY,rand()%GZ}; fiesta_oracle_set_occupancy_vox(O,v,r==0?(rand()%50==0):rand()%2);}
    double*pred=malloc(N*8); memcpy(pred,O->dist,N*8);
    // queue orders come from the oracle's own queues (integration order is validated separately)
    fiesta_oracle_update_occupancy(O,1);
    long nins=fifo_size(&O->q_ins), ndel=fifo_size(&O->q_del);
    u32*ins=malloc(4*(nins+1)),*del=malloc(4*(ndel+1));
    for(long i=0;i<nins;i++){int*v=O->q_ins.e[O->q_ins.head+i].v; ins[i]=vi(v[0],v[1],v[2]);}
    for(long i=0;i<ndel;i++){int*v=O->q_del.e[O->q_del.head+i].v; del[i]=vi(v[0],v[1],v[2]);}
    for(long i=0;i<N;i++) if(pred[i]<0&&O->dist[i]>=0&&C[i]==CU) C[i]=CI;
    expansions=0;
    // E1 insert seeds, in order
    nE=0; cur=0; for(long i=0;i<nins;i++){ long x=ins[i]; if(O->occ[x]>O->l_occ){ int a,b,c; vxyz(x,&a,&b,&c); C[x]=pack(a,b,c); LS[x]=tclock++; E[0][nE++]=x; } }
    // E2 delete: dependants by dense scan, ordered by (delete rank, descending link time)
    u32*rank=malloc(4*N); for(long i=0;i<N;i++)rank[i]=NONE; long nd=0; for(long i=0;i<ndel;i++){ long x=del[i]; if(!(O->occ[x]>O->l_occ) && rank[x]==NONE) rank[x]=nd++; }
    long ndep=0; dep_t*deps=malloc(sizeof(dep_t)*N);
    for(long u=0;u<N;u++){ if(C[u]>=2){ int ox,oy,oz; unpack(C[u],&ox,&oy,&oz); long xo=vi(ox,oy,oz); if(rank[xo]!=NONE){ deps[ndep].k1=rank[xo]; deps[ndep].k2=LS[u]; deps[ndep].v=u; ndep++; } } }
    qsort(deps,ndep,sizeof(dep_t),cmpdep);
    u32*ord=malloc(4*N); for(long i=0;i<N;i++)ord[i]=NONE; for(long i=0;i<ndep;i++)ord[deps[i].v]=i;
    u32*nc0=malloc(4*(ndep+1)),*nc1=malloc(4*(ndep+1)); for(long i=0;i<ndep;i++)nc0[i]=CI;
    for(int it=0;;it++){ long ch=0;
      for(long i=0;i<ndep;i++){ long u=deps[i].v; int x,y,z; vxyz(u,&x,&y,&z); u32 res=CI;
        for(int k=0;k<24;k++){ int nx=x+DIRS[k][0],ny=y+DIRS[k][1],nz=z+DIRS[k][2]; if(!ing(nx,ny,nz)||!inb(nx,ny,nz)) continue; long n=vi(nx,ny,nz); u32 c;
          if(ord[n]!=NONE){ if(ord[n]<(u32)i) c=nc0[ord[n]]; else continue; } else c=C[n];
          if(c>=2 && existc(c)){ res=c; break; } }
        nc1[i]=res; if(res!=nc0[i]) ch++; }
      u32*t=nc0;nc0=nc1;nc1=t; if(!ch) break; }
    for(long i=0;i<ndep;i++){ long u=deps[i].v; C[u]=nc0[i]; LS[u]=tclock++; if(nc0[i]>=2){ E[0][nE++]=u; } }
    fiesta_oracle_update_esdf(O);
    relax();
    long dm=0,cm=0; for(long i=0;i<N;i++){ u32 c=C[i]; double d; int x,y,z; vxyz(i,&x,&y,&z); if(c==CU)d=-10000; else if(c==CI)d=10000; else d=sqrt((double)d2(x,y,z,c))*0.1;
      if(d!=O->dist[i])dm++; int ox=-10000,oy=-10000,oz=-10000; if(c>=2)unpack(c,&ox,&oy,&oz); if(ox!=O->cobs[3*i]||oy!=O->cobs[3*i+1]||oz!=O->cobs[3*i+2])cm++; }
    printf("[gens %ld rounds %ld max %ld evals %ld visits %ld passes %ld | gathers %ld] round %d ins %ld del %ld dep %ld | ref expansions %ld ours %ld | dist mismatches %ld cobs mismatches %ld\n",totgens,totrounds,maxrounds,totevals,totvisits,totpasses,n_gather,r,nins,ndel,ndep,O->st_exp,expansions,dm,cm);
    if(dm||cm||O->st_exp!=expansions) bad++;
    free(pred);free(ins);free(del);free(rank);free(deps);free(ord);free(nc0);free(nc1);
  }
  printf(bad?"FAIL\n":"OK\n");
  return bad?1:0;
}
