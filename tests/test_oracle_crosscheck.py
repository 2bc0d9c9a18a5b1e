"""Cross-check vectors from SURVEY.md Appendix B (produced by an independent fp32 restatement that passes
all 24 reference KAT combos).  NOT reference ground truth: they cover paths the reference's own tests do
not pin (perceptual weights, masks, punch-through, 7-point alpha, NaN axis, single colour).  Two
independently written restatements agreeing is the evidence."""
import numpy as np
from tests import oracle_lib as O

NOISE = bytes.fromhex("D4D38C81DBF50FC5AA8CDFAD085D1B6211457381B5478890D2DDA1B56F0F1382"
                      "A4F01E2245D40E587EA59740D4F98030786751B22F3D040F302A0026B65BE5B5")
P, U = O.PERCEPTUAL, O.UNIFORM
R, C, I = 0, 1, 2
CASES = []


def enc(fmt, px, alg, w, mask=0xFFFF, awa=False):
    return bytes(O.compress_block_masked(fmt, px, mask, O.make_params(alg, w, awa))).hex().upper()


def chk(name, got, exp):
    exp = exp.replace(" ", "")
    CASES.append((name, got, exp))


def build_cases():
    noise = np.frombuffer(NOISE, np.uint8).copy()
    noiseFF = noise.copy(); noiseFF[3::4] = 255
    chk("noise bc1 range P", enc(0,noise,R,P), "8268F4D6E518FFBE")
    chk("noise bc1 cluster P", enc(0,noise,C,P), "CE48B0DEE518FFBE")
    chk("noise bc1 iter P", enc(0,noise,I,P), "CE48B0DEE518FFBE")
    chk("noise bc2 range P", enc(1,noise,R,P), "C86A888B52341AB2D0D74031F24D28D7")
    chk("noise bc2 cluster P", enc(1,noise,C,P), "91CEA43140492A97")
    chk("noise bc2 iter P", enc(1,noise,I,P), "EFC68739604D2897")
    chk("noise bc3 alpha", enc(2,noise,C,P)[:16], "C50F84CA89AFAF5C")
    chk("noise bc4", enc(3,noise,C,P), "DB08C012A133C15F")
    chk("noise bc5", enc(4,noise,C,P), "DB08C012A133C15FF90F026D2B10D1DF")
    chk("noiseFF bc1 range P", enc(0,noiseFF,R,P), "D0D74031F24D28D7")
    chk("noiseFF bc1 range U", enc(0,noiseFF,R,U), "D0D74031684B2E97")
    chk("noiseFF bc1 cluster P", enc(0,noiseFF,C,P), "91CEA43140492A97")
    chk("noiseFF bc1 cluster U", enc(0,noiseFF,C,U), "92CE043240492E97")
    chk("noiseFF bc1 iter P", enc(0,noiseFF,I,P), "A631B0CE15186982")
    chk("noiseFF bc1 iter U", enc(0,noiseFF,I,U), "92CE043240492E97")
    chk("noise bc3 awa cluster P", enc(2,noise,C,P,awa=True), "C50F84CA89AFAF5C72CEE639404D2E97")
    chk("noise bc3 awa iter U", enc(2,noise,I,U,awa=True), "F4CD633248492E17")
    chk("noise m0033 bc3 cluster P", enc(2,noise,C,P,mask=0x33), "81C5080001000000EAE67039A0ADAAAA")
    chk("noiseFF m0033 bc1", enc(0,noiseFF,C,P,mask=0x33), "EAE67039A0ADAAAA")
    chk("noiseFF m0001 bc1", enc(0,noiseFF,C,P,mask=1), "FBFFECBDABAAAAAA")
    chk("noise m0001 bc3", enc(2,noise,C,P,mask=1), "8186000000000000FBFFECBDABAAAAAA")
    grad = np.array([[16*x+8*y,100+4*x,200-8*y,255] for y in range(4) for x in range(4)],np.uint8).reshape(-1)
    chk("grad bc1 range P", enc(0,grad,R,P), "954B3803BDAD2D2F")
    chk("grad bc1 cluster P", enc(0,grad,C,P), "9643380BB5AD2D2B")
    chk("grad bc1 iter P", enc(0,grad,I,P), "9643380BB5AD2D2B")
    chk("grad bc3 range U", enc(2,grad,R,U), "0005FFFFFFFFFFFF954B3803BDAD2F2B")
    chk("grad bc3 cluster U", enc(2,grad,C,U), "7643380BB5AD2F0B")
    chk("grad bc3 iter U", enc(2,grad,I,U), "7643380BB5AD2F0B")
    single = np.array([123,45,67,255]*16,np.uint8)
    chk("single bc1", enc(0,single,C,P), "D9FB4038FFFFFFFF")
    chk("single bc3", enc(2,single,C,P), "0005FFFFFFFFFFFFD9FB4038FFFFFFFF")
    s2 = single.copy()
    for i in range(16):
        if i%3==0: s2[4*i+3]=0
    chk("single a0 bc1 cluster P", enc(0,s2,C,P), "00006879D7755DD7")
    vs=[0,255]+list(range(10,150,10))
    a7=np.array([[v,255-v,0,255] for v in vs],np.uint8).reshape(-1)
    chk("a7 bc4", enc(3,a7,C,P), "0A8C3E206D23D926")
    chk("a7 bc5", enc(4,a7,C,P), "73F577D292DC2601")
    vb=[3,9,17,26,33,41,52,60,66,75,83,92,99,108,117,125]
    a7b=np.array([[v,0,0,v] for v in vb],np.uint8).reshape(-1)
    chk("a7b bc4", enc(3,a7b,C,P), "7D03C96FB7E42601")
    chk("a7b bc3 cluster U", enc(2,a7b,C,U), "7D03C96FB7E4260100680008D5BF2A00")
    rg=np.array([[255,0,0,255] if i%2==0 else [0,255,0,255] for i in range(16)],np.uint8).reshape(-1)
    for a in (R,C,I): chk("redgreen bc1 U %d"%a, enc(0,rg,a,U), "00F800F800000000")
    cols=[[255,0,0,255],[0,255,0,255],[0,0,255,255],[255,0,0,255]]
    rgbr=np.array(cols*4,np.uint8).reshape(-1)
    chk("rgbr range", enc(0,rgbr,R,U), "00F800F800000000")
    chk("rgbr cluster", enc(0,rgbr,C,U), "00F800F882828282")
    chk("rgbr iter", enc(0,rgbr,I,U), "00F800F882828282")
    chk("allmasked bc1", enc(0,noise,C,P,mask=0), "00000000FFFFFFFF")
    chk("allmasked bc3", enc(2,noise,C,P,mask=0), "00050000000000000000000000000000")
    chk("allmasked bc4", enc(3,noise,C,P,mask=0), "0005000000000000")
    z=noise.copy(); z[3::4]=0
    chk("all a0 bc1", enc(0,z,C,P), "00000000FFFFFFFF")


def test_appendix_b_vectors():
    build_cases()
    assert len(CASES) == 44
    bad = [(n, g, e) for n, g, e in CASES if not (g.endswith(e) if len(e) < len(g) else g == e)]
    assert not bad, bad
