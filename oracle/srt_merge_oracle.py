"""TEST INFRASTRUCTURE -- CPU oracle (numpy restatement) of the dubbing tool's timeline merge, the step right after the
vocoder in the reference's SRT pipeline:

    srt_dubbing/src/audio_processor.py:133-230   AudioProcessor._time_synchronized_merge
    srt_dubbing/src/audio_processor.py:70-131    AudioProcessor._natural_concatenation
    srt_dubbing/src/config.py:20-21              AUDIO.DYNAMIC_BUFFER_SIZE = 1024, AUDIO.MAX_AMPLITUDE = 1.0

Only tests/ may import this module.  It is pinned to the unmodified reference class by tests/golden/srt_merge.npz
(oracle/gen_golden.py, group "merge").  `plan_time_synchronized` is the placement arithmetic alone (where each segment
lands and how long the timeline is); the product's host code (b200vgan/timeline.py) is checked against it and the GPU
kernel against `time_synchronized_merge`, bit for bit (fp32 adds in segment order, fp32 division by the peak)."""
import numpy as np

DYNAMIC_BUFFER_SIZE = 1024
MAX_AMPLITUDE = 1.0


def plan_time_synchronized(start_times, end_times, lengths, sample_rate, truncate_on_overflow):
    """audio_processor.py:157-212.  Returns (order, start_sample per sorted segment (-1 = empty, skipped), total_samples)."""
    order = sorted(range(len(start_times)), key=lambda i: start_times[i])     # :157 (stable, like list.sort)
    max_end = 0.0
    for i in order:                                                          # :166-174
        if lengths[i] > 0:
            max_end = max(max_end, start_times[i] + lengths[i] / sample_rate)
        else:
            max_end = max(max_end, end_times[i])
    total = int(max_end * sample_rate) + DYNAMIC_BUFFER_SIZE                 # :176
    starts = []
    for k, i in enumerate(order):
        if lengths[i] == 0:                                                  # :189-192
            starts.append(-1)
            continue
        s = int(start_times[i] * sample_rate)                                # :182
        e = s + lengths[i]
        if not truncate_on_overflow and k > 0:                               # :198-207 (the PREVIOUS segment's nominal start)
            j = order[k - 1]
            prev_end = int(start_times[j] * sample_rate) + lengths[j]
            if s < prev_end:
                s = prev_end
                e = s + lengths[i]
        if e > total:                                                        # :210-218 (the array grows)
            total = e + DYNAMIC_BUFFER_SIZE
        starts.append(s)
    return order, starts, total


def time_synchronized_merge(segments, sample_rate, truncate_on_overflow):
    """segments: list of dicts with 'start_time', 'end_time' (optional) and 'audio_data' (1-D float32)."""
    st = [float(s["start_time"]) for s in segments]
    en = [float(s.get("end_time", s["start_time"])) for s in segments]
    audio = [np.asarray(s["audio_data"], dtype=np.float32) for s in segments]
    order, starts, total = plan_time_synchronized(st, en, [len(a) for a in audio], sample_rate, truncate_on_overflow)
    merged = np.zeros(total, dtype=np.float32)
    for i, s in zip(order, starts):
        if s >= 0:
            merged[s:s + len(audio[i])] += audio[i]                          # :221
    if not truncate_on_overflow:                                             # :227-232
        mx = np.max(np.abs(merged))
        if mx > MAX_AMPLITUDE:
            merged = merged / mx
    return merged


def natural_concatenation(segments):
    """audio_processor.py:86-127: segments in subtitle-index order, empty ones skipped, back to back."""
    parts = [np.asarray(s["audio_data"], dtype=np.float32) for s in sorted(segments, key=lambda x: x.get("index", 0))]
    parts = [p for p in parts if len(p)]
    return np.concatenate(parts) if parts else np.array([])
