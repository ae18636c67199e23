"""Drop-in mirror of the reference command line (reference processAXCTD.py:47-183):
same flags, same settings dict, same output-file format.

Two extra, optional flags that the reference does not have:
  --wired     make -s/-e/-a/-b/-p/-t/-l/-u act as documented (as shipped they are
              inert or crash); default is faithful behaviour
  --device N  CUDA device index
"""
from __future__ import annotations

import argparse
import logging
import os

from . import AXCTDprocessor


# flag, long name, default, help -- the flags and defaults of the reference CLI (processAXCTD.py:52-66); the help
# texts are this package's own
_VALUE_FLAGS = (
    ("-i", "--input", "ERROR_NO_FILE_SPECIFIED", "WAV recording to decode"),
    ("-o", "--output", "output.txt", "text file the profile is written to"),
    ("-s", "--starttime", "0", "where the drop starts in the recording (seconds, or [HH:]MM:SS)"),
    ("-e", "--endtime", "-1", "where the drop ends in the recording (-1: end of file)"),
    ("-a", "--autodetect-start", "30", "earliest profile start the 7500 Hz detector accepts, seconds after the first 400 Hz pulse"),
    ("-b", "--autodetect-end", "-1", "latest profile start: the profile is taken to begin here if no 7500 Hz tone was seen (-1: never)"),
    ("-p", "--sig-threshold-400", "2", "400 Hz level (log10 ratio to the dead frequency) that marks a header pulse"),
    ("-t", "--sig-threshold-7500", "1.5", "rise of the 7500 Hz level that marks the profile start"),
    ("-d", "--dead-freq", "3000", "quiet reference frequency of the signal levels (Hz)"),
    ("-l", "--pointsperloop", "100000", "samples per processing iteration"),
    ("-m", "--mark-freq", "400", "frequency of a 1 bit (Hz)"),
    ("-n", "--space-freq", "800", "frequency of a 0 bit (Hz)"),
)


def main(argv=None):
    parser = argparse.ArgumentParser(description="AXCTD audio recording -> profile text file (B200 engine)")
    for short, long_name, default, text in _VALUE_FLAGS:
        parser.add_argument(short, long_name, default=default, help=text)
    parser.add_argument("-u", "--use-bandpass", action="store_true", help="band-pass 100-1200 Hz instead of the 1200 Hz low-pass")
    parser.add_argument("--wired", action="store_true", help="make the documented flags act (not reference behaviour)")
    parser.add_argument("--device", type=int, default=0, help="CUDA device index")
    args = parser.parse_args(argv)

    # processAXCTD.py:71-76: the two messages and the silent exit are part of the interface
    if args.input == "ERROR_NO_FILE_SPECIFIED":
        print("[!] Error- no input WAV file specified! Terminating")
        exit()
    elif not os.path.exists(args.input):
        print("[!] Specified input file does not exist! Terminating")
        exit()

    def window(lo_text, hi_text):                # :80-91: a non-positive end means "open"
        lo, hi = parse_times(lo_text), parse_times(hi_text)
        return [lo, hi if hi > 0 else -1]

    timerange = window(args.starttime, args.endtime)
    triggerrange = window(args.autodetect_start, args.autodetect_end)
    settings = {"triggerrange": triggerrange,                                # keys and types of :93-99
                "minR400": float(args.sig_threshold_400),
                "mindR7500": float(args.sig_threshold_7500),
                "deadfreq": float(args.dead_freq),
                "pointsperloop": int(args.pointsperloop),
                "mark_space_freqs": [float(args.mark_freq), float(args.space_freq)],
                "use_bandpass": args.use_bandpass}
    return processAXCTD(args.input, args.output, timerange, settings,
                        mode="wired" if args.wired else "faithful", device=args.device)


def parse_times(time_string):
    """Seconds from "SS", "MM:SS" or "HH:MM:SS" (same results as processAXCTD.py:106-121: fields beyond the hours are
    dropped with a log message, anything that is not an integer gives -2)."""
    fields = time_string.split(":")
    if len(fields) > 3:
        logging.info("[!] time fields beyond HH:MM:SS are ignored")
    try:
        values = [int(f) for f in fields[::-1][:3]]          # seconds, minutes, hours; the ignored fields are not looked at
    except ValueError:
        logging.info("[!] time not understood, taking -2 (treated as unset)")
        return -2
    return sum(v * 60 ** i for i, v in enumerate(values))


def wired_settings(settings, f_s):
    """CLI key names -> the processor's internal names (what the flags document)."""
    return {'minr400': settings['minR400'], 'mindr7500': settings['mindR7500'], 'deadfreq': settings['deadfreq'],
            'refreshrate': settings['pointsperloop'] / f_s, 'mark_space_freqs': settings['mark_space_freqs'],
            'usebandpass': settings['use_bandpass']}


def profile_lines(ap, wavfile, timerange, settings, defaults=None):
    """Lines of the output file in order, processAXCTD.py:146-183.  ``defaults``
    supplies the '(default)' coefficient lines; without it an incomplete header
    raises KeyError('zcoeff_default') at the same point as the reference
    (:161-167), after the lines before it were produced."""
    minR400 = settings['minR400']
    mindR7500 = settings['mindR7500']
    deadfreq = settings['deadfreq']
    pointsperloop = settings['pointsperloop']
    triggerrange = settings['triggerrange']
    yield f"AXCTD profile for {wavfile}\n"
    fs = ap.f_s
    yield f'Sampling frequency (fs): {fs} Hz\n'
    yield f'Audio file length: {ap.numpoints/fs} sec\n'
    yield f'400 Hz pulse start: {ap.firstpulse400/fs} sec\n'
    yield f'7500 Hz tone start: {ap.profstartind/fs} sec\n'
    yield "\nAXCTD header information:\n"
    for desc, ckey in zip(['Probe Code', 'Maximum Depth (m)', 'Probe Serial'], ['probe_code', 'max_depth', 'serial_no']):
        yield f"{desc}: {ap.metadata[ckey]}\n"
    yield "Conversion equations:\n"
    md = dict(ap.metadata)
    if defaults:
        md.update(defaults)
    for coeff, desc, symb in zip(['z', 't', 'c'], ['Depth', 'Temperature', 'Conductivity'], ['t', 'T', 'C']):
        if sum(ap.metadata[coeff + 'coeff_valid']) == 4:
            cfield = coeff + 'coeff'
            defaultstatus = ''
        else:
            cfield = coeff + 'coeff_default'
            defaultstatus = '(default)'
        cureqn = ' + '.join([f'{val}*{symb}^{i}' for i, val in enumerate(md[cfield])])
        yield f'{desc}: {cureqn} {defaultstatus}\n'
    yield '\nProcessor Settings:\n'
    yield f'Time Range: {timerange[0]} sec to {timerange[1] if timerange[1] >= 0 else "N/A"} sec\n'
    yield f'Min. 400 Hz power ratio: {minR400}\n'
    yield f'Min. 7500 Hz power ratio: {mindR7500}\n'
    yield f'Dead frequency: {deadfreq}\n'
    yield f'Points per loop: {pointsperloop}\n'
    yield f'Trigger range: {triggerrange[0]} sec to {triggerrange[1] if triggerrange[1] >= 0 else "N/A"} sec\n'
    yield '\nAXCTD Profile:\n'
    yield 'Time (s), Hex Frame, Depth (m), Temperature (C), Conductivity (mS/cm), Salinity (PSU)\n'
    for (t, hf, z, T, C, S) in zip(ap.time, ap.hexframes, ap.depth, ap.temperature, ap.conductivity, ap.salinity):
        yield f"{t:8.2f},  {hf},{z:10.2f},{T:16.2f},{C:21.2f},{S:15.2f}\n"


def format_profile(ap, wavfile, timerange, settings, defaults=None):
    return "".join(profile_lines(ap, wavfile, timerange, settings, defaults))


def processAXCTD(wavfile, outfile, timerange, settings, mode="faithful", device=0, engine=None):
    """processAXCTD.py:126-183."""
    for key in ('minR400', 'mindR7500', 'deadfreq', 'pointsperloop', 'triggerrange', 'mark_space_freqs', 'use_bandpass'):
        settings[key]                                                        # KeyError like :128-134
    print("Processing profile")
    if mode == "wired":
        probe_fs = AXCTDprocessor.read_wav_pcm16(wavfile)[0]
        ap = AXCTDprocessor.AXCTD_Processor(wavfile, timerange=timerange, user_settings=wired_settings(settings, probe_fs),
                                            mode="wired", device=device, engine=engine)
        ap.triggerrange = list(settings['triggerrange'])
        defaults = {c + 'coeff_default': ap.settings[c + 'coeff_axctd'] for c in 'ztc'}
    else:
        ap = AXCTDprocessor.AXCTD_Processor(wavfile, timerange=timerange, user_settings=settings, device=device, engine=engine)
        defaults = None
    ap.run()
    print("Profile processing complete- writing output files")
    with open(outfile, 'w') as f:
        # line by line, so that a KeyError leaves the same partial file the reference leaves
        for line in profile_lines(ap, wavfile, timerange, settings, defaults):
            f.write(line)


if __name__ == "__main__":
    main()
