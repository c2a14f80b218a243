//! GPU blocks with radiorust's own block API.
//!
//! Each block is the reference block's Tokio task loop (`filters.rs:169-273`, `transform.rs:306-362`,
//! `resampling.rs:64-141`, `modulation.rs:105-147`) with the DSP replaced by one `rr_chain_push`; constructors
//! and setters carry the reference's names and argument meaning.  They are made `Consumer` / `Producer` by
//! radiorust's `impl_block_trait!` (src/blocks/mod.rs:111-144), so `feed_into` / `feed_from` work unchanged.
//!
//! | reference | here |
//! |---|---|
//! | `blocks::FreqShifter<Flt>` | [`GpuFreqShifter<Flt>`] |
//! | `blocks::filters::Filter<Flt>` | [`GpuFilter<Flt>`] |
//! | `blocks::Downsampler<Flt>` / `Upsampler<Flt>` | [`GpuDownsampler<Flt>`] / [`GpuUpsampler<Flt>`] |
//! | `blocks::modulation::FmDemod<Flt>` | [`GpuFmDemod<Flt>`] |
//! | a `feed_into` chain of the above | [`GpuChain<Flt>`]: one block, samples stay on the device between stages |
//!
//! Like the reference blocks they must be created inside a Tokio runtime (they `spawn`), and dropping one does not
//! stop its task (src/blocks/mod.rs:27-34).  Without a B200 the constructors panic: there is no CPU fallback.
use crate::chain::{Chain, FreqResp, Stage, Window};
use crate::pool::PinnedChunkBufPool;
use crate::{Context, GpuFloat};

use num::Complex;
use radiorust::bufferpool::ChunkBufPool;
use radiorust::flow::*;
use radiorust::impl_block_trait;
use radiorust::signal::*;
use tokio::sync::{mpsc, watch};
use tokio::task::spawn;

/// Parameter changes travel to the block task and are applied at chunk granularity, like the reference's
/// `watch::Receiver::has_changed()` polls (filters.rs:179, transform.rs:318, modulation.rs:113)
enum Control {
    SetShift { stage: usize, shift: f64 },
    UpdateFilter { stage: usize, freq_resp: FreqResp, window: Option<Window> },
    SetDeviation { stage: usize, deviation: f64 },
    SetGain { stage: usize, gain: f64 },
}

fn apply<Flt: GpuFloat>(chain: &mut Chain<Flt>, c: Control) {
    let r = match c {
        Control::SetShift { stage, shift } => chain.set_shift(stage, None, shift),
        Control::UpdateFilter { stage, freq_resp, window } => chain.update_filter(stage, freq_resp, window),
        Control::SetDeviation { stage, deviation } => chain.set_deviation(stage, deviation),
        Control::SetGain { stage, gain } => chain.set_gain(stage, gain),
    };
    // contract violations panic inside the task, like the reference's assert!s (resampling.rs:51-56, 77-83)
    r.unwrap_or_else(|e| panic!("{e}"));
}

/// The `SamplesLost` event of `radiorust::blocks::chunks::events` is what a Rechunker / Overlapper stage inside a
/// chain reports; re-exported here so that the task can send it ahead of the event that caused it.
use radiorust::blocks::chunks::events::SamplesLost;

/// Spawns the block task: recv -> (staging copy into pinned memory) -> `rr_chain_push` -> send.
fn spawn_task<Flt: GpuFloat>(
    mut chain: Chain<Flt>,
    mut receiver: Receiver<Signal<Complex<Flt>>>,
    sender: Sender<Signal<Complex<Flt>>>,
    mut control: mpsc::UnboundedReceiver<Control>,
    output_chunk_len: Option<usize>,
) {
    let mut buf_pool = ChunkBufPool::<Complex<Flt>>::new();
    let mut pinned = PinnedChunkBufPool::<Complex<Flt>>::new(chain.context()).unwrap_or_else(|e| panic!("{e}"));
    spawn(async move {
        loop {
            let Ok(signal) = receiver.recv().await else { return; };
            while let Ok(c) = control.try_recv() {
                apply(&mut chain, c);
            }
            match signal {
                Signal::Samples { sample_rate, chunk: input_chunk } => {
                    let n = input_chunk.len();
                    if n == 0 {
                        continue;
                    }
                    let lost_before = chain.samples_lost_count();
                    let cap = chain.max_output(sample_rate, n, 1).max(1);
                    // chain edges: one memcpy into / out of pinned memory, asynchronous DMA in between
                    let mut pin_in = pinned.get_with_capacity(n).unwrap_or_else(|e| panic!("{e}"));
                    pin_in.extend_from_slice(&input_chunk);
                    let pin_in = pin_in.finalize();
                    let mut pin_out = pinned.get_with_capacity(cap).unwrap_or_else(|e| panic!("{e}"));
                    let pushed = chain.push_pinned(sample_rate, &pin_in, &mut pin_out).unwrap_or_else(|e| panic!("{e}"));
                    // a Rechunker stage that met a new sample rate reports the dropped partial chunk first (chunks.rs:71-79)
                    for _ in lost_before..chain.samples_lost_count() {
                        let Ok(()) = sender.send(Signal::new_event(SamplesLost)).await else { return; };
                    }
                    // resamplers emit whole chunks of output_chunk_len (resampling.rs:121-131); every other block
                    // answers one chunk with at most one chunk
                    let piece_len = output_chunk_len.unwrap_or(pushed.count.max(1));
                    for piece in pin_out[..pushed.count].chunks(piece_len) {
                        let mut out = buf_pool.get_with_capacity(piece.len());
                        out.extend_from_slice(piece);
                        let Ok(()) = sender.send(Signal::Samples { sample_rate: pushed.sample_rate, chunk: out.finalize() }).await
                        else { return; };
                    }
                }
                Signal::Event(event) => {
                    // Filter history / FmDemod's previous sample are dropped on interrupts (filters.rs:262-267,
                    // modulation.rs:133-138); resamplers and the NCO keep their state
                    let lost = chain.event(event.is_interrupt()).unwrap_or_else(|e| panic!("{e}"));
                    for _ in 0..lost {
                        let Ok(()) = sender.send(Signal::new_event(SamplesLost)).await else { return; };
                    }
                    let Ok(()) = sender.send(Signal::Event(event)).await else { return; };
                }
            }
        }
    });
}

fn open(device: i32) -> Context {
    Context::new(device).unwrap_or_else(|e| panic!("{e}"))
}

fn single<Flt: GpuFloat>(device: i32, stage: Stage) -> Chain<Flt> {
    Chain::new(&open(device), vec![stage], 1).unwrap_or_else(|e| panic!("{e}"))
}

// -------------------------------------------------------------------------------------------------------------
// FreqShifter (transform.rs:266-391)
// -------------------------------------------------------------------------------------------------------------
/// GPU version of `blocks::FreqShifter`
pub struct GpuFreqShifter<Flt> {
    receiver_connector: ReceiverConnector<Signal<Complex<Flt>>>,
    sender_connector: SenderConnector<Signal<Complex<Flt>>>,
    precision: f64,
    shift: watch::Sender<f64>,
    control: mpsc::UnboundedSender<Control>,
}
impl_block_trait! { <Flt> Consumer<Signal<Complex<Flt>>> for GpuFreqShifter<Flt> }
impl_block_trait! { <Flt> Producer<Signal<Complex<Flt>>> for GpuFreqShifter<Flt> }

impl<Flt: GpuFloat> GpuFreqShifter<Flt> {
    /// 1 Hz precision, zero shift (transform.rs:282-284)
    pub fn new(device: i32) -> Self {
        Self::with_precision_and_shift(device, 1.0, 0.0)
    }
    /// 1 Hz precision and initial `shift` in hertz (transform.rs:287-289)
    pub fn with_shift(device: i32, shift: f64) -> Self {
        Self::with_precision_and_shift(device, 1.0, shift)
    }
    /// given `precision` in hertz, zero shift (transform.rs:292-294)
    pub fn with_precision(device: i32, precision: f64) -> Self {
        Self::with_precision_and_shift(device, precision, 0.0)
    }
    /// given `precision` and `shift`, both in hertz (transform.rs:297)
    pub fn with_precision_and_shift(device: i32, precision: f64, shift: f64) -> Self {
        let (receiver, receiver_connector) = new_receiver::<Signal<Complex<Flt>>>();
        let (sender, sender_connector) = new_sender::<Signal<Complex<Flt>>>();
        let (control, control_rx) = mpsc::unbounded_channel();
        let chain = single::<Flt>(device, Stage::FreqShifter { precision, shift });
        spawn_task(chain, receiver, sender, control_rx, None);
        Self { receiver_connector, sender_connector, precision, shift: watch::channel(shift).0, control }
    }
    /// Frequency precision in hertz (fixed at creation, transform.rs:376-378)
    pub fn precision(&self) -> f64 {
        self.precision
    }
    /// Current frequency shift (transform.rs:380-382)
    pub fn shift(&self) -> f64 {
        *self.shift.borrow()
    }
    /// Set frequency shift; phase continuous, effective at the next chunk (transform.rs:384-386, 322-327)
    pub fn set_shift(&self, shift: f64) {
        self.shift.send_replace(shift);
        self.control.send(Control::SetShift { stage: 0, shift }).ok();
    }
    /// Update frequency shift (transform.rs:388-390)
    pub fn update_shift<F: FnOnce(&mut f64)>(&self, modify: F) {
        self.shift.send_modify(modify);
        self.control.send(Control::SetShift { stage: 0, shift: self.shift() }).ok();
    }
}

// -------------------------------------------------------------------------------------------------------------
// Filter (filters.rs:110-298)
// -------------------------------------------------------------------------------------------------------------
/// GPU version of `blocks::filters::Filter`: fast convolution with a designed frequency response, one chunk of
/// delay, history dropped on redesign and on interrupt events
pub struct GpuFilter<Flt> {
    receiver_connector: ReceiverConnector<Signal<Complex<Flt>>>,
    sender_connector: SenderConnector<Signal<Complex<Flt>>>,
    control: mpsc::UnboundedSender<Control>,
}
impl_block_trait! { <Flt> Consumer<Signal<Complex<Flt>>> for GpuFilter<Flt> }
impl_block_trait! { <Flt> Producer<Signal<Complex<Flt>>> for GpuFilter<Flt> }

impl<Flt: GpuFloat> GpuFilter<Flt> {
    /// Kaiser window with its first null at bin 2.0 (filters.rs:128-133)
    pub fn new<F>(device: i32, freq_resp: F) -> Self
    where
        F: Fn(isize, f64) -> Complex<f64> + Send + Sync + 'static,
    {
        Self::new_internal(device, Box::new(freq_resp), Window::kaiser_with_null_at_bin(2.0))
    }
    /// Rectangular window (filters.rs:138-143)
    pub fn new_rectangular<F>(device: i32, freq_resp: F) -> Self
    where
        F: Fn(isize, f64) -> Complex<f64> + Send + Sync + 'static,
    {
        Self::new_internal(device, Box::new(freq_resp), Window::Rectangular)
    }
    /// Given window function (filters.rs:145-152)
    pub fn with_window<F, W>(device: i32, freq_resp: F, window: W) -> Self
    where
        F: Fn(isize, f64) -> Complex<f64> + Send + Sync + 'static,
        W: radiorust::windowing::Window + Send + Sync + 'static,
    {
        Self::new_internal(device, Box::new(freq_resp), Window::Custom(Box::new(window)))
    }
    fn new_internal(device: i32, freq_resp: FreqResp, window: Window) -> Self {
        let (receiver, receiver_connector) = new_receiver::<Signal<Complex<Flt>>>();
        let (sender, sender_connector) = new_sender::<Signal<Complex<Flt>>>();
        let (control, control_rx) = mpsc::unbounded_channel();
        let chain = single::<Flt>(device, Stage::Filter { freq_resp, window });
        spawn_task(chain, receiver, sender, control_rx, None);
        Self { receiver_connector, sender_connector, control }
    }
    /// Update the frequency response, keep the window (filters.rs:279-286)
    pub fn update<F>(&self, freq_resp: F)
    where
        F: Fn(isize, f64) -> Complex<f64> + Send + Sync + 'static,
    {
        self.control.send(Control::UpdateFilter { stage: 0, freq_resp: Box::new(freq_resp), window: None }).ok();
    }
    /// Update frequency response and window (filters.rs:288-297)
    pub fn update_with_window<F, W>(&self, freq_resp: F, window: W)
    where
        F: Fn(isize, f64) -> Complex<f64> + Send + Sync + 'static,
        W: radiorust::windowing::Window + Send + Sync + 'static,
    {
        self.control
            .send(Control::UpdateFilter { stage: 0, freq_resp: Box::new(freq_resp), window: Some(Window::Custom(Box::new(window))) })
            .ok();
    }
}

// -------------------------------------------------------------------------------------------------------------
// Downsampler / Upsampler (resampling.rs)
// -------------------------------------------------------------------------------------------------------------
macro_rules! resampler_block {
    ($name:ident, $stage:ident, $doc:literal) => {
        #[doc = $doc]
        pub struct $name<Flt> {
            receiver_connector: ReceiverConnector<Signal<Complex<Flt>>>,
            sender_connector: SenderConnector<Signal<Complex<Flt>>>,
        }
        impl_block_trait! { <Flt> Consumer<Signal<Complex<Flt>>> for $name<Flt> }
        impl_block_trait! { <Flt> Producer<Signal<Complex<Flt>>> for $name<Flt> }

        impl<Flt: GpuFloat> $name<Flt> {
            /// `quality` 3.0 (resampling.rs:38-40 / :173-175)
            pub fn new(device: i32, output_chunk_len: usize, output_rate: f64, bandwidth: f64) -> Self {
                Self::with_quality(device, output_chunk_len, output_rate, bandwidth, 3.0)
            }
            /// With `quality` >= 1.0; panics on the argument errors the reference asserts (resampling.rs:51-56 / :186-188)
            pub fn with_quality(device: i32, output_chunk_len: usize, output_rate: f64, bandwidth: f64, quality: f64) -> Self {
                let (receiver, receiver_connector) = new_receiver::<Signal<Complex<Flt>>>();
                let (sender, sender_connector) = new_sender::<Signal<Complex<Flt>>>();
                let (_control, control_rx) = mpsc::unbounded_channel();
                let chain = single::<Flt>(device, Stage::$stage { output_chunk_len, output_rate, bandwidth, quality });
                spawn_task(chain, receiver, sender, control_rx, Some(output_chunk_len));
                Self { receiver_connector, sender_connector }
            }
        }
    };
}
resampler_block!(GpuDownsampler, Downsampler, "GPU version of `blocks::Downsampler` (resampling.rs:14-146)");
resampler_block!(GpuUpsampler, Upsampler, "GPU version of `blocks::Upsampler` (resampling.rs:149-280)");

// -------------------------------------------------------------------------------------------------------------
// FmDemod (modulation.rs:83-158)
// -------------------------------------------------------------------------------------------------------------
/// GPU version of `blocks::modulation::FmDemod`
pub struct GpuFmDemod<Flt> {
    receiver_connector: ReceiverConnector<Signal<Complex<Flt>>>,
    sender_connector: SenderConnector<Signal<Complex<Flt>>>,
    deviation: watch::Sender<f64>,
    control: mpsc::UnboundedSender<Control>,
}
impl_block_trait! { <Flt> Consumer<Signal<Complex<Flt>>> for GpuFmDemod<Flt> }
impl_block_trait! { <Flt> Producer<Signal<Complex<Flt>>> for GpuFmDemod<Flt> }

impl<Flt: GpuFloat> GpuFmDemod<Flt> {
    /// FM demodulator with given frequency deviation in hertz (modulation.rs:97)
    pub fn new(device: i32, deviation: f64) -> Self {
        let (receiver, receiver_connector) = new_receiver::<Signal<Complex<Flt>>>();
        let (sender, sender_connector) = new_sender::<Signal<Complex<Flt>>>();
        let (control, control_rx) = mpsc::unbounded_channel();
        let chain = single::<Flt>(device, Stage::FmDemod { deviation });
        spawn_task(chain, receiver, sender, control_rx, None);
        Self { receiver_connector, sender_connector, deviation: watch::channel(deviation).0, control }
    }
    /// Frequency deviation in hertz (modulation.rs:150-152)
    pub fn deviation(&self) -> f64 {
        *self.deviation.borrow()
    }
    /// Set frequency deviation in hertz (modulation.rs:154-157)
    pub fn set_deviation(&self, deviation: f64) -> &Self {
        self.deviation.send_replace(deviation);
        self.control.send(Control::SetDeviation { stage: 0, deviation }).ok();
        self
    }
}

// -------------------------------------------------------------------------------------------------------------
// Fused chain: several reference blocks as ONE block
// -------------------------------------------------------------------------------------------------------------
/// `FreqShifter -> Filter -> Downsampler` (or any other sequence of the blocks above, plus `GainControl`) as one
/// block: samples cross PCIe once per direction and the stages run as fused kernels.  Built with
/// [`GpuChain::builder`]; setters address a stage by its position in the chain.
pub struct GpuChain<Flt> {
    receiver_connector: ReceiverConnector<Signal<Complex<Flt>>>,
    sender_connector: SenderConnector<Signal<Complex<Flt>>>,
    control: mpsc::UnboundedSender<Control>,
}
impl_block_trait! { <Flt> Consumer<Signal<Complex<Flt>>> for GpuChain<Flt> }
impl_block_trait! { <Flt> Producer<Signal<Complex<Flt>>> for GpuChain<Flt> }

/// Builder of a [`GpuChain`]
pub struct GpuChainBuilder<Flt> {
    device: i32,
    stages: Vec<Stage>,
    output_chunk_len: Option<usize>,
    _flt: std::marker::PhantomData<Flt>,
}

impl<Flt: GpuFloat> GpuChainBuilder<Flt> {
    /// append `FreqShifter::with_precision_and_shift(precision, shift)`
    pub fn freq_shifter(mut self, precision: f64, shift: f64) -> Self {
        self.stages.push(Stage::FreqShifter { precision, shift });
        self
    }
    /// append `Filter::new(freq_resp)`
    pub fn filter<F>(mut self, freq_resp: F) -> Self
    where
        F: Fn(isize, f64) -> Complex<f64> + Send + Sync + 'static,
    {
        self.stages.push(Stage::Filter { freq_resp: Box::new(freq_resp), window: Window::kaiser_with_null_at_bin(2.0) });
        self
    }
    /// append `Filter::new_rectangular(freq_resp)` (the de-emphasis filter of examples/relm_app/simple_receiver.rs:43-49)
    pub fn filter_rectangular<F>(mut self, freq_resp: F) -> Self
    where
        F: Fn(isize, f64) -> Complex<f64> + Send + Sync + 'static,
    {
        self.stages.push(Stage::Filter { freq_resp: Box::new(freq_resp), window: Window::Rectangular });
        self
    }
    /// append `Downsampler::new(output_chunk_len, output_rate, bandwidth)`
    pub fn downsampler(mut self, output_chunk_len: usize, output_rate: f64, bandwidth: f64) -> Self {
        self.stages.push(Stage::Downsampler { output_chunk_len, output_rate, bandwidth, quality: 3.0 });
        self.output_chunk_len = Some(output_chunk_len);
        self
    }
    /// append `Upsampler::new(output_chunk_len, output_rate, bandwidth)`
    pub fn upsampler(mut self, output_chunk_len: usize, output_rate: f64, bandwidth: f64) -> Self {
        self.stages.push(Stage::Upsampler { output_chunk_len, output_rate, bandwidth, quality: 3.0 });
        self.output_chunk_len = Some(output_chunk_len);
        self
    }
    /// append `FmDemod::new(deviation)`
    pub fn fm_demod(mut self, deviation: f64) -> Self {
        self.stages.push(Stage::FmDemod { deviation });
        self.output_chunk_len = None;
        self
    }
    /// append `GainControl::new(gain)`
    pub fn gain_control(mut self, gain: f64) -> Self {
        self.stages.push(Stage::GainControl { gain });
        self
    }
    /// append any other stage
    pub fn stage(mut self, stage: Stage) -> Self {
        self.stages.push(stage);
        self
    }
    /// Create the block (spawns its task; panics without a B200)
    pub fn build(self) -> GpuChain<Flt> {
        let (receiver, receiver_connector) = new_receiver::<Signal<Complex<Flt>>>();
        let (sender, sender_connector) = new_sender::<Signal<Complex<Flt>>>();
        let (control, control_rx) = mpsc::unbounded_channel();
        let chain = Chain::<Flt>::new(&open(self.device), self.stages, 1).unwrap_or_else(|e| panic!("{e}"));
        spawn_task(chain, receiver, sender, control_rx, self.output_chunk_len);
        GpuChain { receiver_connector, sender_connector, control }
    }
}

impl<Flt: GpuFloat> GpuChain<Flt> {
    /// Start a chain on CUDA device `device`
    pub fn builder(device: i32) -> GpuChainBuilder<Flt> {
        GpuChainBuilder { device, stages: Vec::new(), output_chunk_len: None, _flt: std::marker::PhantomData }
    }
    /// `FreqShifter::set_shift` of the stage at position `stage`
    pub fn set_shift(&self, stage: usize, shift: f64) {
        self.control.send(Control::SetShift { stage, shift }).ok();
    }
    /// `Filter::update` of the stage at position `stage`
    pub fn update_filter<F>(&self, stage: usize, freq_resp: F)
    where
        F: Fn(isize, f64) -> Complex<f64> + Send + Sync + 'static,
    {
        self.control.send(Control::UpdateFilter { stage, freq_resp: Box::new(freq_resp), window: None }).ok();
    }
    /// `FmDemod::set_deviation` of the stage at position `stage`
    pub fn set_deviation(&self, stage: usize, deviation: f64) {
        self.control.send(Control::SetDeviation { stage, deviation }).ok();
    }
    /// `GainControl::set` of the stage at position `stage`
    pub fn set_gain(&self, stage: usize, gain: f64) {
        self.control.send(Control::SetGain { stage, gain }).ok();
    }
}
