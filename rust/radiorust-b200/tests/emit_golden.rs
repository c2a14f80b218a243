//! Pins the repository's oracle to REAL radiorust.
//!
//! `tests/golden/*.npy` in the repository are outputs of the numpy oracle (no Rust toolchain exists where it is
//! developed, and radiorust ships no test vectors for Filter / FreqShifter / Downsampler / Upsampler / FmDemod:
//! their test modules are empty, filters.rs:378, resampling.rs:282, modulation.rs:160).  On a machine with cargo:
//!
//! ```text
//! python tests/golden/make_golden.py --emit-inputs /tmp/rr_golden      # inputs + case list + oracle outputs
//! RR_GOLDEN_DIR=/tmp/rr_golden cargo test --release --test emit_golden -- --nocapture
//! ```
//!
//! runs the unmodified radiorust blocks (`feed_into` chains, Tokio runtime) on the same inputs, writes
//! `<case>.radiorust.bin` next to them and compares with the oracle's `<case>.oracle.bin`: relative L2 error
//! <= 1e-5 (f32) / 1e-12 (f64).  No GPU is needed for this test; `gpu_blocks_match_radiorust` below runs the same
//! cases through the GPU blocks when a B200 is present (`RR_GPU_DEVICE=0`).
//!
//! File formats (written by make_golden.py): `cases.txt` has one line per case:
//! `name flt sample_rate chunk_len n_chunks block;block;...` with blocks
//! `freqshift:precision:shift`, `filter_lowpass:cutoff`, `filter_deemph:tau`, `downsample:len:rate:bw:quality`,
//! `upsample:len:rate:bw:quality`, `fmdemod:deviation`; `<case>.input.bin` / `.oracle.bin` are interleaved
//! little-endian (re, im) pairs of f32 or f64.
use num::Complex;
use radiorust::blocks::filters::{deemphasis_factor, Filter};
use radiorust::blocks::modulation::FmDemod;
use radiorust::blocks::{Downsampler, FreqShifter, Upsampler};
use radiorust::bufferpool::Chunk;
use radiorust::flow::*;
use radiorust::numbers::Float;
use radiorust::signal::Signal;

use std::path::{Path, PathBuf};
use std::time::Duration;

#[derive(Clone, Debug)]
struct Case {
    name: String,
    flt: String,
    sample_rate: f64,
    chunk_len: usize,
    n_chunks: usize,
    blocks: Vec<Vec<String>>,
}

fn golden_dir() -> Option<PathBuf> {
    std::env::var("RR_GOLDEN_DIR").ok().map(PathBuf::from)
}

fn read_cases(dir: &Path) -> Vec<Case> {
    let text = std::fs::read_to_string(dir.join("cases.txt")).expect("cases.txt (run make_golden.py --emit-inputs)");
    text.lines()
        .filter(|l| !l.trim().is_empty() && !l.starts_with('#'))
        .map(|l| {
            let f: Vec<&str> = l.split_whitespace().collect();
            Case {
                name: f[0].into(),
                flt: f[1].into(),
                sample_rate: f[2].parse().unwrap(),
                chunk_len: f[3].parse().unwrap(),
                n_chunks: f[4].parse().unwrap(),
                blocks: f[5].split(';').map(|b| b.split(':').map(str::to_owned).collect()).collect(),
            }
        })
        .collect()
}

trait Sample: Float + Send + Sync + 'static {
    fn from_le(bytes: &[u8]) -> Self;
    fn to_le(self, out: &mut Vec<u8>);
    fn as_f64(self) -> f64;
    const SIZE: usize;
    const TOL: f64;
}
impl Sample for f32 {
    fn from_le(b: &[u8]) -> Self {
        f32::from_le_bytes([b[0], b[1], b[2], b[3]])
    }
    fn to_le(self, out: &mut Vec<u8>) {
        out.extend_from_slice(&self.to_le_bytes());
    }
    fn as_f64(self) -> f64 {
        self as f64
    }
    const SIZE: usize = 4;
    const TOL: f64 = 1e-5;
}
impl Sample for f64 {
    fn from_le(b: &[u8]) -> Self {
        f64::from_le_bytes([b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7]])
    }
    fn to_le(self, out: &mut Vec<u8>) {
        out.extend_from_slice(&self.to_le_bytes());
    }
    fn as_f64(self) -> f64 {
        self
    }
    const SIZE: usize = 8;
    const TOL: f64 = 1e-12;
}

fn read_complex<Flt: Sample>(path: &Path) -> Vec<Complex<Flt>> {
    let bytes = std::fs::read(path).unwrap_or_else(|e| panic!("{}: {e}", path.display()));
    bytes.chunks_exact(2 * Flt::SIZE).map(|c| Complex::new(Flt::from_le(&c[..Flt::SIZE]), Flt::from_le(&c[Flt::SIZE..]))).collect()
}

fn write_complex<Flt: Sample>(path: &Path, data: &[Complex<Flt>]) {
    let mut out = Vec::with_capacity(data.len() * 2 * Flt::SIZE);
    for v in data {
        v.re.to_le(&mut out);
        v.im.to_le(&mut out);
    }
    std::fs::write(path, out).unwrap();
}

fn rel_l2<Flt: Sample>(a: &[Complex<Flt>], b: &[Complex<Flt>]) -> f64 {
    assert_eq!(a.len(), b.len(), "output lengths differ");
    let (mut num, mut den) = (0.0f64, 0.0f64);
    for (x, y) in a.iter().zip(b) {
        let (dr, di) = (x.re.as_f64() - y.re.as_f64(), x.im.as_f64() - y.im.as_f64());
        num += dr * dr + di * di;
        den += y.re.as_f64() * y.re.as_f64() + y.im.as_f64() * y.im.as_f64();
    }
    (num / den).sqrt()
}

fn lowpass(cutoff: f64) -> impl Fn(isize, f64) -> Complex<f64> + Send + Sync + 'static {
    move |_, freq| if freq.abs() <= cutoff { Complex::from(1.0) } else { Complex::from(0.0) }
}
fn deemph(tau: f64) -> impl Fn(isize, f64) -> Complex<f64> + Send + Sync + 'static {
    // examples/relm_app/simple_receiver.rs:43-49
    move |bin, freq| if bin != 0 && freq.abs() >= 20.0 && freq.abs() <= 16000.0 { deemphasis_factor(tau, freq) } else { Complex::from(0.0) }
}

/// Keeps the blocks of a chain alive while it runs
enum AnyBlock<Flt> {
    Shift(FreqShifter<Flt>),
    Filter(Filter<Flt>),
    Down(Downsampler<Flt>),
    Up(Upsampler<Flt>),
    Fm(FmDemod<Flt>),
}

/// Runs `case` through real radiorust blocks; returns the concatenated output samples
async fn run_reference<Flt: Sample>(case: &Case, input: &[Complex<Flt>]) -> Vec<Complex<Flt>> {
    let (sender, sender_connector) = new_sender::<Signal<Complex<Flt>>>();
    let (mut receiver, receiver_connector) = new_receiver::<Signal<Complex<Flt>>>();
    let mut blocks: Vec<AnyBlock<Flt>> = Vec::new();
    for b in &case.blocks {
        let p = |i: usize| -> f64 { b[i].parse().unwrap() };
        let blk = match b[0].as_str() {
            "freqshift" => AnyBlock::Shift(FreqShifter::<Flt>::with_precision_and_shift(p(1), p(2))),
            "filter_lowpass" => AnyBlock::Filter(Filter::<Flt>::new(lowpass(p(1)))),
            "filter_deemph" => AnyBlock::Filter(Filter::<Flt>::new_rectangular(deemph(p(1)))),
            "downsample" => AnyBlock::Down(Downsampler::<Flt>::with_quality(p(1) as usize, p(2), p(3), p(4))),
            "upsample" => AnyBlock::Up(Upsampler::<Flt>::with_quality(p(1) as usize, p(2), p(3), p(4))),
            "fmdemod" => AnyBlock::Fm(FmDemod::<Flt>::new(p(1))),
            other => panic!("unknown block {other}"),
        };
        blocks.push(blk);
    }
    // wire: sender -> blocks[0] -> ... -> receiver
    fn consumer<Flt: Float>(b: &AnyBlock<Flt>) -> &ReceiverConnector<Signal<Complex<Flt>>> {
        match b {
            AnyBlock::Shift(x) => x.receiver_connector(),
            AnyBlock::Filter(x) => x.receiver_connector(),
            AnyBlock::Down(x) => x.receiver_connector(),
            AnyBlock::Up(x) => x.receiver_connector(),
            AnyBlock::Fm(x) => x.receiver_connector(),
        }
    }
    fn producer<Flt: Float>(b: &AnyBlock<Flt>) -> &SenderConnector<Signal<Complex<Flt>>> {
        match b {
            AnyBlock::Shift(x) => x.sender_connector(),
            AnyBlock::Filter(x) => x.sender_connector(),
            AnyBlock::Down(x) => x.sender_connector(),
            AnyBlock::Up(x) => x.sender_connector(),
            AnyBlock::Fm(x) => x.sender_connector(),
        }
    }
    consumer(&blocks[0]).connect(&sender_connector);
    for i in 1..blocks.len() {
        consumer(&blocks[i]).connect(producer(&blocks[i - 1]));
    }
    receiver_connector.connect(producer(blocks.last().unwrap()));

    let chunks: Vec<Vec<Complex<Flt>>> = input.chunks(case.chunk_len).map(|c| c.to_vec()).collect();
    let sample_rate = case.sample_rate;
    let feeder = tokio::spawn(async move {
        for c in chunks {
            if sender.send(Signal::Samples { sample_rate, chunk: Chunk::from(c) }).await.is_err() {
                return;
            }
        }
        // keep the sender alive until the consumer side has drained
        tokio::time::sleep(Duration::from_secs(3600)).await;
    });
    let mut out = Vec::new();
    // the chain is drained when nothing arrives for a while (blocks never close their outputs on their own)
    loop {
        match tokio::time::timeout(Duration::from_secs(5), receiver.recv()).await {
            Ok(Ok(Signal::Samples { chunk, .. })) => out.extend_from_slice(&chunk),
            Ok(Ok(Signal::Event(_))) => {}
            Ok(Err(_)) | Err(_) => break,
        }
    }
    feeder.abort();
    out
}

async fn pin_case<Flt: Sample>(dir: &Path, case: &Case) {
    let input = read_complex::<Flt>(&dir.join(format!("{}.input.bin", case.name)));
    assert_eq!(input.len(), case.chunk_len * case.n_chunks);
    let got = run_reference::<Flt>(case, &input).await;
    write_complex(&dir.join(format!("{}.radiorust.bin", case.name)), &got);
    let want = read_complex::<Flt>(&dir.join(format!("{}.oracle.bin", case.name)));
    let err = rel_l2(&want, &got);
    println!("{:20} {} samples: oracle vs radiorust rel_l2 = {:.3e} (tolerance {:.0e})", case.name, got.len(), err, Flt::TOL);
    assert!(err <= Flt::TOL, "{}: the oracle differs from radiorust by {err:e}", case.name);
}

#[tokio::test(flavor = "multi_thread")]
async fn oracle_matches_radiorust() {
    let Some(dir) = golden_dir() else {
        eprintln!("RR_GOLDEN_DIR not set: skipped (see the header of this file)");
        return;
    };
    for case in read_cases(&dir) {
        match case.flt.as_str() {
            "f32" => pin_case::<f32>(&dir, &case).await,
            "f64" => pin_case::<f64>(&dir, &case).await,
            other => panic!("unknown float type {other}"),
        }
    }
}

/// The same cases through the GPU blocks (needs a B200: set RR_GPU_DEVICE)
#[tokio::test(flavor = "multi_thread")]
async fn gpu_blocks_match_radiorust() {
    let (Some(dir), Ok(dev)) = (golden_dir(), std::env::var("RR_GPU_DEVICE")) else {
        eprintln!("RR_GOLDEN_DIR / RR_GPU_DEVICE not set: skipped");
        return;
    };
    let device: i32 = dev.parse().unwrap();
    use radiorust_b200::blocks::GpuChain;
    for case in read_cases(&dir).into_iter().filter(|c| c.flt == "f32") {
        let mut b = GpuChain::<f32>::builder(device);
        for blk in &case.blocks {
            let p = |i: usize| -> f64 { blk[i].parse().unwrap() };
            b = match blk[0].as_str() {
                "freqshift" => b.freq_shifter(p(1), p(2)),
                "filter_lowpass" => b.filter(lowpass(p(1))),
                "filter_deemph" => b.filter_rectangular(deemph(p(1))),
                "downsample" => b.downsampler(p(1) as usize, p(2), p(3)),
                "upsample" => b.upsampler(p(1) as usize, p(2), p(3)),
                "fmdemod" => b.fm_demod(p(1)),
                other => panic!("unknown block {other}"),
            };
        }
        let gpu = b.build();
        let (sender, sender_connector) = new_sender::<Signal<Complex<f32>>>();
        let (mut receiver, receiver_connector) = new_receiver::<Signal<Complex<f32>>>();
        gpu.feed_from(&sender_connector);
        gpu.feed_into(&receiver_connector);
        let input = read_complex::<f32>(&dir.join(format!("{}.input.bin", case.name)));
        let chunks: Vec<Vec<Complex<f32>>> = input.chunks(case.chunk_len).map(|c| c.to_vec()).collect();
        let sample_rate = case.sample_rate;
        let feeder = tokio::spawn(async move {
            for c in chunks {
                if sender.send(Signal::Samples { sample_rate, chunk: Chunk::from(c) }).await.is_err() {
                    return;
                }
            }
            tokio::time::sleep(Duration::from_secs(3600)).await;
        });
        let mut got = Vec::new();
        loop {
            match tokio::time::timeout(Duration::from_secs(5), receiver.recv()).await {
                Ok(Ok(Signal::Samples { chunk, .. })) => got.extend_from_slice(&chunk),
                Ok(Ok(Signal::Event(_))) => {}
                Ok(Err(_)) | Err(_) => break,
            }
        }
        feeder.abort();
        let want = read_complex::<f32>(&dir.join(format!("{}.radiorust.bin", case.name)));
        let err = rel_l2(&got, &want);
        println!("{:20} GPU blocks vs radiorust rel_l2 = {err:.3e}", case.name);
        assert!(err <= 1e-5, "{}: GPU blocks differ from radiorust by {err:e}", case.name);
    }
}
