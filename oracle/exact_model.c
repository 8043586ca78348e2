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
//             the summaries of its targets are recomputed.  Short work lists are evaluated by gathering while last round's
//             flips are refreshed concurrently; long ones first bring the summaries up to date (targets of the flips, or
//             everything when most entries flipped) and then evaluate through them.  The fixpoint is reached by a round
//             without flips: it has read final words only.
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
static sum_t *SUM; static u32 *SUMg;
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
static u32 *W[3]; static long nW[3]; static u32 *F[3]; static long nF[3]; static u32 *wstamp; static u32 wclock=0;
static void shuffle(u32*a,long n){ for(long i=n-1;i>0;i--){ long j=rand()%(i+1); u32 t=a[i];a[i]=a[j];a[j]=t; } }
static long expansions; static long totrounds=0, totgens=0, maxrounds=0, totevals=0, totdense=0;
static u32 *emask; static u32 *slotc;   // slotc: SMALL mode only
static long DENSE_MIN=16;               // work lists longer than this are evaluated through the summaries (kernel: 4096)
static u32 sclock=0;
// summaries of the targets of element i; every target is summarised once per pass (stamp sclock in SUMg)
static void claim_summaries(long i){ long p=E[cur][i]; int x,y,z; vxyz(p,&x,&y,&z);
  for(int k=0;k<25;k++){ int nx=k<24?x+DIRS[k][0]:x,ny=k<24?y+DIRS[k][1]:y,nz=k<24?z+DIRS[k][2]:z; if(!ing(nx,ny,nz)||(k<24&&!inb(nx,ny,nz))) continue;
    long n=vi(nx,ny,nz); if(SUMg[n]!=sclock){ SUMg[n]=sclock; summarize(nx,ny,nz); } } }
// after element i flipped: its offer (i,k) carries timestamp 32i+k whatever its code, so the summary of target k changes only if
// that timestamp is its `first` / `best`, or if the element's new offer would become one of them
static void refresh_targets(long i){ long p=E[cur][i]; int x,y,z; vxyz(p,&x,&y,&z); u64 w=MB[p];
  for(int k=0;k<25;k++){ int nx=k<24?x+DIRS[k][0]:x,ny=k<24?y+DIRS[k][1]:y,nz=k<24?z+DIRS[k][2]:z; if(!ing(nx,ny,nz)||(k<24&&!inb(nx,ny,nz))) continue;
    sum_t u=SUM[vi(nx,ny,nz)]; if(u.snap_c==CU) continue; u32 ts=(u32)i*32+k; int redo=u.first==ts||u.best_ts==ts;
    if(!redo){ u64 kind=mb_kind(w); u32 c=mb_code(w);
      if(((kind==K_PUSH&&k<24)||(kind==K_PULL&&k==24))&&c>=2){ u32 d=d2(nx,ny,nz,c);
        if(u.best_ts==NONE) redo=1; else { u32 bd=d2(nx,ny,nz,u.best_c); redo=ts<u.first||d<bd||(d==bd&&ts<u.best_ts); } } }
    if(redo) summarize(nx,ny,nz); } }
static void relax(void){
  int big=nE>SMALL;
  for(long i=0;i<nE;i++) MB[E[cur][i]]=mbw((u32)i,K_PUSH,C[E[cur][i]]);   // initial guess: everybody pushes its snapshot code
  while(nE){
    totgens++;
    if(big){ sclock++; u32*ord=malloc(4*(nE+1)); for(long i=0;i<nE;i++)ord[i]=(u32)i; shuffle(ord,nE); for(long q=0;q<nE;q++) claim_summaries(ord[q]); free(ord); }
    int rounds=0; nW[0]=nW[1]=nW[2]=0; nF[0]=nF[1]=nF[2]=0;
    for(int r=1;;r++){
      int in=r%3,out=(r+1)%3; nW[(r+2)%3]=0; nF[(r+2)%3]=0; wclock++;
      long nw; u32*wl=W[in];
      if(r==1){ nw=nE; for(long i=0;i<nE;i++)wl[i]=(u32)i; } else nw=nW[in];
      long nf=(big&&r>1)?nF[in]:0;
      if(r>1&&nw==0&&nf==0) break;
      rounds++;
      int dense=big&&r>1&&nw>DENSE_MIN; int use_sum=big&&(r==1||dense);
      if(dense){ totdense++;     // summaries first (targets of last round's flips, or everything), then evaluate through them
        if(nf<nE/4){ for(long q=0;q<nf;q++) refresh_targets(F[in][q]); } else { sclock++; for(long i=0;i<nE;i++) claim_summaries(i); } }
      // this round's work: evaluations and (big, not dense) summary refreshes of last round's flips, randomly interleaved
      long nref=(big&&!dense)?nf:0; long tot=nw+nref; u32*ord=malloc(4*(tot+1)); for(long q=0;q<tot;q++)ord[q]=(u32)q; shuffle(ord,tot);
      for(long q=0;q<tot;q++){ long a=ord[q];
        if(a>=nw){ refresh_targets(F[in][a-nw]); continue; }
        long i=wl[a]; totevals++;
        u64 nb=eval(i,use_sum); long p=E[cur][i];
        if(nb!=MB[p]){ int wide=mb_kind(MB[p])==K_PUSH||mb_kind(nb)==K_PUSH;   // only a pushing element reaches beyond its own voxel: 129 offsets, else {0} u dirs
          MB[p]=nb; if(big) F[out][nF[out]++]=(u32)i;
          int x,y,z; vxyz(p,&x,&y,&z);
          for(int o=0;o<(wide?nOFF:25);o++){ int nx=x+OFF[o][0],ny=y+OFF[o][1],nz=z+OFF[o][2]; if(!ing(nx,ny,nz)) continue; u64 w=MB[vi(nx,ny,nz)]; if(w==MB_NONE) continue; u32 j=mb_idx(w);
            if(j>(u32)i && wstamp[j]!=wclock){ wstamp[j]=wclock; W[out][nW[out]++]=j; } } } }
      free(ord);
      if(rounds>100000){printf("no convergence\n");exit(1);} }
    totrounds+=rounds; if(rounds>maxrounds)maxrounds=rounds;
    // commit: slot masks
    long total=0;
    for(long i=0;i<nE;i++){ u64 b=MB[E[cur][i]]; u32 m=0; long p=E[cur][i]; int x,y,z; vxyz(p,&x,&y,&z);
      if(mb_kind(b)!=K_DEAD) expansions++;
      if(mb_kind(b)==K_PUSH){ for(int k=0;k<24;k++){ int nx=x+DIRS[k][0],ny=y+DIRS[k][1],nz=z+DIRS[k][2]; if(!ing(nx,ny,nz)||!inb(nx,ny,nz)) continue; u32 ts=(u32)i*32+k;
          if(big){ sum_t u=SUM[vi(nx,ny,nz)]; if(u.best_ts==ts) m|=1u<<k; } else { st_t f=gather(nx,ny,nz,NONE,NULL); if(f.ts==ts){ m|=1u<<k; slotc[i*32+k]=f.c; } } } }
      else if(mb_kind(b)==K_PULL){ u32 ts=(u32)i*32+24; if(big){ sum_t u=SUM[p]; if(u.best_ts==ts) m|=1u<<24; } else { st_t f=gather(x,y,z,NONE,NULL); if(f.ts==ts){ m|=1u<<24; slotc[i*32+24]=f.c; } } }
      emask[i]=m; total+=__builtin_popcount(m); }
    for(long i=0;i<nE;i++) MB[E[cur][i]]=MB_NONE;    // (kernel: retired by compare-and-swap in the apply phase)
    // apply: exclusive scan of the mask popcounts -> positions
    int big2=total>SMALL; long r=0;
    for(long i=0;i<nE;i++){ long p=E[cur][i]; int x,y,z; vxyz(p,&x,&y,&z); u32 m=emask[i];
      for(int k=0;k<25;k++) if(m>>k&1){ int nx=k<24?x+DIRS[k][0]:x,ny=k<24?y+DIRS[k][1]:y,nz=k<24?z+DIRS[k][2]:z; long v=vi(nx,ny,nz);
          u32 c=big?SUM[v].best_c:slotc[i*32+k]; C[v]=c; LS[v]=tclock+(u64)i*32+k; E[cur^1][r]=(u32)v; MB[v]=mbw((u32)r,K_PUSH,c); r++; } }
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
  if(argc>8) DENSE_MIN=atol(argv[8]);
  SUM=calloc(N,sizeof(sum_t)); SUMg=calloc(N,4); emask=malloc(4*N); slotc=malloc(4*32*(SMALL+1));
  E[0]=malloc(4*N);E[1]=malloc(4*N); for(int q=0;q<3;q++){W[q]=malloc(4*N);F[q]=malloc(4*N);} wstamp=calloc(N,4);
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
    printf("[gens %ld rounds %ld max %ld evals %ld dense %ld | queries %ld gathers %ld] round %d ins %ld del %ld dep %ld | ref expansions %ld ours %ld | dist mismatches %ld cobs mismatches %ld\n",totgens,totrounds,maxrounds,totevals,totdense,n_query,n_gather,r,nins,ndel,ndep,O->st_exp,expansions,dm,cm);
    if(dm||cm||O->st_exp!=expansions) bad++;
    free(pred);free(ins);free(del);free(rank);free(deps);free(ord);free(nc0);free(nc1);
  }
  printf(bad?"FAIL\n":"OK\n");
  return bad?1:0;
}
